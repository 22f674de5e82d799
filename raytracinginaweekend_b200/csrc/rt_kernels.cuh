// rt_kernels.cuh — the wavefront path tracer kernels (sm_100a).
//
//   k_generate : Size2i::iterf + pixel jitter + Camera::ray   (size2i.rs:38-56, rendering.rs:174-175, camera.rs:175-200)
//   k_extend   : Scene::hit -> BVH closest hit                 (hittable.rs:429-473)          [dominant kernel]
//   k_shade    : one iteration of ray_color's loop + Material::scatter + regeneration + compaction
//                                                             (rendering.rs:19-92, material.rs, texture.rs, background_color.rs)
//   k_finalize : per-pixel mean + merge_planes                 (rendering.rs:178-179, 239-252)
//   k_trace    : Scene::hit on caller rays (validation entry)
#pragma once
#include "rt_device.cuh"

namespace rt {

struct Counters {
    uint32_t nb[2];         // live continuing (bounce) paths in queue 0 / 1: positions [0, nb)
    uint32_t nn[2];         // live regenerated (primary) paths in queue 0 / 1: positions cap-1 .. cap-nn (descending)
    uint32_t cursor_extend; // dynamic work fetch
    uint32_t pad;
    unsigned long long next_chain;  // next (lane, pixel) chain of this pipeline to hand to a slot whose chain has ended
    uint32_t n_susp[2];     // suspended traversals waiting in pool 0 / 1 (may over-count: readers clamp to RenderDev::susp_cap)
    uint32_t budget;        // rounds a warp may run after it can no longer refill before it parks its rays (adapted by k_shade)
    uint32_t boost;         // log2 of the factor on top of that budget while too large a share of the rays gets parked
    unsigned long long tail_acc;  // this iteration's queue launch: (warps that ran after they could no longer refill) << 40 | (sum of those rounds)
    uint32_t pad4;
    uint32_t main_warps_left;  // warps of the queue launch of k_extend that have not finished yet (set by k_shade / k_generate)
    unsigned long long rays;   // closest-hit queries so far
    unsigned long long paths;  // camera samples generated so far
    unsigned long long node_visits, prim_tests, exact_steps;  // only maintained by the COUNT kernel variant
    unsigned long long rounds, busy_lanes, tail_rounds, tail_busy_lanes;  // COUNT: warp rounds / busy lanes, overall and after the queue ran dry
    unsigned long long ref_rounds, ref_lanes;  // COUNT: warp rounds in which >= 1 lane traced on the reference tree, and those lanes
    unsigned long long redo_total;     // row f4 (COUNT variant): rays the fast traversal could not prove -> reference traversal
    uint32_t max_sp, pad2;             // COUNT variant: deepest traversal stack use of any ray
    unsigned long long hist[24];       // COUNT variant: rays by floor(log2(rounds the ray took)) — the per-launch tail
};

struct RenderDev {
    float4* A[2];   // origin.xyz, bits(pixel)           (the lane it renders is D.y; with it the accumulator = acc[lane][pixel])
    float4* B[2];   // direction.xyz, bits(sample k within the lane)
    float4* C[2];   // throughput.rgb, primary direction.y (Q6)
    int2* D[2];     // remaining depth, local lane m the path belongs to
    float* Tm[2];   // Ray::time of the path (camera.rs:185-190); NULL when the scene has no Animation leaf
    uint2* hits;    // (child ref of closest primitive | RT_NO_HIT, bits(t)) per queue position
    float4* acc;    // [n_local_lanes * npix] running per-lane sums (rgb) + sum of r^2+g^2+b^2 per sample (w);
                    // exclusive owner: the single live path of the track that currently renders that lane
    Counters* ctr;
    uint32_t cap;            // queue capacity = number of path slots of THIS pipeline
    unsigned long long n_chains;  // chains of this pipeline.  A chain = all samples of one (lane, pixel), rendered one after the other by
                             // ONE slot (so every per-lane sum keeps the reference's order); chain c = (pixel c % npix, the pipeline's
                             // (c / npix)-th lane).  A slot whose chain has ended takes the next chain from Counters::next_chain, so
                             // every wavefront stays full until the chains run out, whatever the paths' lengths.
    uint32_t* iter_log;      // optional [max_iters]: live paths of each iteration (debug / per-launch accounting)
    float* long_rays;        // optional (COUNT + RT_LONG_RAYS): [0] = count (as u32 bits), then 8 floats per ray that took >= 4096 rounds:
                             // origin, direction, rounds, flags (1 = finished on the reference tree)
    uint32_t iter;           // iteration index of this launch (only used with iter_log)
    uint32_t pipe, n_pipes;  // rank-local lanes are dealt round-robin to n_pipes independent wavefront pipelines (streams)
    // Suspended traversals.  A launch cannot end before its longest ray has; a handful of rays per frame take 10^3 - 10^4 rounds (a
    // sphere 1800 units away whose test has lost its digits to cancellation accepts hits metres off, so nothing near it can be
    // culled; hundreds of coplanar fan triangles) and the whole GPU waits for them, every iteration.  Once a warp can no longer
    // refill, it gives its rays `tail_budget` more rounds and then parks what is still running: (cursor, stack, best hit) go to
    // pool q^1, the hit slot says RT_SUSPENDED, k_shade puts the path back into the queue untouched and links the pool entry to its
    // new position, and the next extend resumes the parked traversals FIRST, so they run alongside a whole launch of ordinary rays.
    uint2* susp[2];             // pools: entry = 4 header words-pairs (pos, cur | sp, best | end, dpos | -, -) + stack_cap stack entries
    uint32_t susp_cap, susp_stride;  // entries per pool, uint2 per entry
    uint32_t tail_budget;       // rounds a warp keeps going after it can no longer refill (0 = never suspend)
    uint32_t tail_fixed;        // 1: the budget is tail_budget itself, not adapted (rt_set_option("tail_budget", -n): tests that force parking)
    uint32_t extend_warps;      // warps of the queue launch of k_extend (Counters::main_warps_left starts from it every iteration)
    uint32_t refill_threshold;  // extend refills idle lanes when fewer than this many lanes are busy
    uint32_t leaf_weight;       // phase vote: leaf phase runs when n_leaf * leaf_weight > n_node * 16
    uint32_t npix, width, height;
    int32_t max_depth;
    uint32_t n_tasks;        // T' = lanes that got >= 1 sample (split_work_tasks)
    uint32_t lane_first, lane_stride, n_local_lanes;
    uint32_t spp_whole, spp_rem;  // lane id gets spp_whole + (id < spp_rem)
    uint32_t seed0, seed1;
    float sx, sy;            // 1/(w-1), 1/(h-1)   (size2i.rs:21-22; also the jitter scale, rendering.rs:132-138)
    uint32_t mode;
    uint32_t debug_skip_redo;  // measurement only (rt_set_option("debug_skip_redo")): unproven fast-path results are kept — NOT the reference's result
    uint32_t ref_only;       // traversal MODE 0: the root gate must not apply the tight cull either
};

#define RT_ROOT_MISS 0x80000000u  // top bit of the per-path sample index: the ray fails the BVH root test (producer-side)
#define RT_FORCE_REF 0x40000000u  // next bit: trace this segment on the reference tree (k_shade could not prove the fast-path hit)
#define RT_RESUME 0x20000000u     // next bit: this segment's traversal is parked in the suspend pool (the queue entry itself is skipped)
#define RT_K_MASK 0x1FFFFFFFu     // the sample index itself (make_plan caps the samples per lane at 2^29)
#define RT_TAIL_BASE 0x80000000u  // k_extend: batch_next == batch_end >= this = rounds since the warp could no longer refill (+ this)
#define RT_SUSPENDED 0xFFFFFFFEu  // hits[pos].x: traversal parked, hits[pos].y = pool entry

__device__ __forceinline__ uint32_t lane_samples(const RenderDev& R, uint32_t lane_id) {
    return R.spp_whole + (lane_id < R.spp_rem ? 1u : 0u);
}

// New camera path for (pixel, lane, sample k).
__device__ __forceinline__ void generate_path(const SceneDev& S, const RenderDev& R, uint32_t pixel, uint32_t lane_id,
                                              uint32_t k, float3& o, float3& d, float& time) {
    uint32_t x = pixel % R.width, y = pixel / R.width;
    uint32_t s_global = k * R.n_tasks + lane_id;
    uint4 r = philox4x32_10(R.seed0, R.seed1, pixel, s_global, 0u, 0u);
    float fx = (float)x * R.sx, fy = (float)y * R.sy;          // size2i.rs:52-55
    float px = fx + (u01(r.x) * R.sx + 0.0f);                   // Uniform<Vec2f>: x then y (vec2.rs:77-82)
    float py = fy + (u01(r.y) * R.sy + 0.0f);
    camera_ray(S.cam, px, py, r, o, d);
    // shutter time (camera.rs:185-190): the next draw of the camera block after the lens disc (draw #2 without a lens,
    // draw #4 = first output of sub-block 1 with one)
    time = S.cam.time_start;
    if (S.cam.time_start != S.cam.time_end) {
        uint32_t bits = r.z;
        if (S.cam.lens_radius > 0.0f) bits = philox4x32_10(R.seed0, R.seed1, pixel, s_global, 0u, 1u).x;
        time = S.cam.time_start + u01(bits) * (S.cam.time_end - S.cam.time_start);
    }
    time = time + (S.cam.pace_x * px + S.cam.pace_y * py);  // + Vec2f::dot(shutter_pace, point), camera.rs:190
}

__global__ void __launch_bounds__(256) k_generate(const __grid_constant__ SceneDev S, const __grid_constant__ RenderDev R) {
    const uint32_t n_slots = R.cap;  // slot j starts with chain j: pixel j % npix of the pipeline's (j / npix)-th lane (local lane m = pipe + n_pipes * that)
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n_slots; j += gridDim.x * blockDim.x) {
        const uint32_t pixel = j % R.npix, track = R.pipe + R.n_pipes * (j / R.npix);
        const uint32_t lane_id = R.lane_first + track * R.lane_stride;
        float3 o, d;
        float time;
        generate_path(S, R, pixel, lane_id, 0u, o, d, time);
        if (R.Tm[0]) R.Tm[0][j] = time;
        R.A[0][j] = make_float4(o.x, o.y, o.z, __uint_as_float(pixel));
        const uint32_t miss = root_gate(S, o, d, 0.001f, __int_as_float(0x7f800000), R.ref_only != 0u) ? 0u : RT_ROOT_MISS;
        R.B[0][j] = make_float4(d.x, d.y, d.z, __uint_as_float(miss));
        R.C[0][j] = make_float4(1.0f, 1.0f, 1.0f, d.y);
        R.D[0][j] = make_int2(R.max_depth, (int)track);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        R.ctr->nb[0] = n_slots; R.ctr->nn[0] = 0;
        R.ctr->nb[1] = 0; R.ctr->nn[1] = 0;
        R.ctr->cursor_extend = 0;
        R.ctr->next_chain = n_slots;
        R.ctr->n_susp[0] = 0; R.ctr->n_susp[1] = 0;
        R.ctr->budget = R.tail_fixed ? R.tail_budget : 1u << 24;  // the first launch measures; budgets follow from the tails seen (k_shade)
        R.ctr->boost = 0; R.ctr->tail_acc = 0ull;
        R.ctr->main_warps_left = R.extend_warps;
        R.ctr->rays = 0;
        R.ctr->paths = n_slots;
        R.ctr->node_visits = 0;
        R.ctr->prim_tests = 0;
        R.ctr->exact_steps = 0;
        R.ctr->rounds = 0; R.ctr->busy_lanes = 0; R.ctr->tail_rounds = 0; R.ctr->tail_busy_lanes = 0;
        R.ctr->redo_total = 0; R.ctr->max_sp = 0; R.ctr->ref_rounds = 0; R.ctr->ref_lanes = 0;
        for (int h = 0; h < 24; ++h) R.ctr->hist[h] = 0;
    }
}

// Volume draw key of a path segment in the wavefront (computed on demand: only volume leaves ask for it)
template <int FEAT_>
struct ExtendKey {
    static constexpr int FEAT = FEAT_;
    const RenderDev* R;
    int q;
    uint32_t pos;
    __device__ __forceinline__ VolKey operator()() const {
        const uint32_t pixel = __float_as_uint(__ldcg(&R->A[q][pos].w)), k = __float_as_uint(__ldcg(&R->B[q][pos].w)) & RT_K_MASK;
        const int2 dm = __ldcg(&R->D[q][pos]);
        const uint32_t lane_id = R->lane_first + (uint32_t)dm.y * R->lane_stride;
        return VolKey{R->seed0, R->seed1, pixel, k * R->n_tasks + lane_id, 1u + (uint32_t)(R->max_depth - dm.x)};
    }
};
// rt_trace: keyed by the ray index, seed 0, block 0
struct TraceKey {
    static constexpr int FEAT = 2;
    uint64_t i;
    __device__ __forceinline__ VolKey operator()() const { return VolKey{0u, 0u, (uint32_t)i, (uint32_t)(i >> 32), 0u}; }
};

// Queue position of the i-th live path: continuing paths grow from 0, regenerated ones from the top.
__device__ __forceinline__ uint32_t queue_pos(uint32_t i, uint32_t nb, uint32_t cap) { return i < nb ? i : cap - 1u - (i - nb); }

// Closest hit for every live path of queue `q`.  Persistent warps: every lane owns one ray's traversal state and,
// when fewer than `refill_threshold` lanes are still busy, the idle lanes pull fresh rays with one warp-aggregated
// atomic — rays differ by 100x in node visits, so waiting for the slowest lane would idle most of the warp.
// MODE 3 (row f4): fast traversal of the rebuilt topology + proof that the reference returns the same element; a lane whose
// result is not proven restarts on the reference tree (MODE 2 code) in place, overlapped with the other lanes' work — a
// separate redo launch would serialise the few longest rays of every iteration into its own tail.
#define RT_REFMODE_BIT 16u  // Trav::dpos: this lane is (re)tracing on the reference tree
// Parked traversals (RenderDev::susp): deliberately not inlined — the rare path must not cost the traversal loop registers.
__device__ __noinline__ void park_lane(uint2* e, const uint2* stk, uint32_t stk_stride, uint32_t pos, uint32_t cur, uint32_t sp, uint32_t best,
                                       float end, uint32_t dpos) {
    __stcg(e, make_uint2(pos, cur));  // .x is rewritten by k_shade with the path's position in the next queue
    __stcg(e + 1, make_uint2(sp, best));
    __stcg(e + 2, make_uint2(__float_as_uint(end), dpos));
    for (uint32_t i = 0; i < sp; ++i) __stcg(e + 4 + i, stk[i * stk_stride]);
}
__device__ __noinline__ void unpark_lane(const uint2* e, uint2* stk, uint32_t stk_stride, uint32_t& cur, uint32_t& sp, uint32_t& best, float& end,
                                         uint32_t& dpos) {
    const uint2 h0 = __ldcg(e), h1 = __ldcg(e + 1), h2 = __ldcg(e + 2);
    cur = h0.y; sp = h1.x; best = h1.y; end = __uint_as_float(h2.x); dpos = h2.y;
    for (uint32_t i = 0; i < h1.x; ++i) stk[i * stk_stride] = __ldcg(e + 4 + i);
}
template <int MODE, bool COUNT, int FEAT, bool RESUME>
__global__ void __launch_bounds__(128, 8) k_extend(const __grid_constant__ SceneDev S, const __grid_constant__ RenderDev R, int q) {
    extern __shared__ uint2 smem[];
    uint2* stk = reinterpret_cast<uint2*>(smem) + threadIdx.x;
    // RESUME = false: the queue.  RESUME = true (a second, small launch that runs beside it on its own stream): the traversals parked by
    // the previous extend (pool q); warp w owns entries 32 w .. 32 w + 31, so every entry has a lane without any cursor.
    const uint32_t nb = R.ctr->nb[q];
    const uint32_t n = RESUME ? min(R.ctr->n_susp[q], R.susp_cap) : nb + R.ctr->nn[q];
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t FULL = 0xFFFFFFFFu;
    const float INF = __int_as_float(0x7f800000);
    if (!RESUME && blockIdx.x == 0 && threadIdx.x == 0) { R.ctr->nb[q ^ 1] = 0; R.ctr->nn[q ^ 1] = 0; }  // nobody touches them during extend
    uint32_t resume_rounds = 0;  // RESUME only
    TravCount tc;
    tc.nodes = 0; tc.prims = 0; tc.exact = 0; tc.redo = 0; tc.max_sp = 0;
    Trav T;
    bool active = false, exhausted = false;
    uint32_t my_pos = 0;
    uint32_t batch_next = 0, batch_end = 0;  // warp-uniform: ray indices [batch_next, batch_end) already claimed by this warp
    uint32_t act_mask = 0;                   // warp-uniform ballot of `active`
    unsigned long long rounds_c = 0, busy_c = 0, trounds_c = 0, tbusy_c = 0, rrounds_c = 0, rlanes_c = 0;
    uint32_t my_rounds = 0;  // COUNT: scheduling rounds this lane's current ray has been active for
    for (;;) {  // (outer loop: only taken again when the suspend pool is full)
    for (;;) {
        // refill: idle lanes take rays from the warp's claimed batch (one atomic per 32 rays)
        if (act_mask != FULL && !(exhausted && batch_next == batch_end) && (uint32_t)__popc(act_mask) < R.refill_threshold) {
            if (batch_next == batch_end) {
                uint32_t base = 0;
                if (RESUME) {
                    base = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32u;
                    exhausted = true;  // this warp's one and only batch
                } else {
                    if (lane == 0) base = atomicAdd(&R.ctr->cursor_extend, 32u);
                    base = __shfl_sync(FULL, base, 0);
                    if (base + 32u >= n) exhausted = true;
                }
                batch_next = min(base, n);
                batch_end = min(base + 32u, n);
            }
            const uint32_t idle = ~act_mask;
            const uint32_t avail = batch_end - batch_next;
            const uint32_t rank = __popc(idle & ((1u << lane) - 1u));
            if (!active && rank < avail) {
                const uint2* pe = nullptr;
                if (RESUME) {  // a parked traversal takes its ray from its (relinked) queue position, cursor and stack from the pool
                    pe = R.susp[q] + (size_t)(batch_next + rank) * R.susp_stride;
                    my_pos = __ldcg(pe).x;
                } else {
                    my_pos = queue_pos(batch_next + rank, nb, R.cap);
                }
                const float4 a = R.A[q][my_pos], b = R.B[q][my_pos];
                if (!RESUME && (__float_as_uint(b.w) & (RT_ROOT_MISS | RT_RESUME))) {
                    // RT_RESUME: parked, the RESUME launch finishes this segment; otherwise the ray fails the root test
                    if (!(__float_as_uint(b.w) & RT_RESUME)) R.hits[my_pos] = make_uint2(RT_NO_HIT, __float_as_uint(INF));
                } else {
                    trav_setup(S, T, f3(a.x, a.y, a.z), f3(b.x, b.y, b.z), R.Tm[q] ? R.Tm[q][my_pos] : 0.0f, 0.001f, INF);
                    active = true;
                    if (COUNT) my_rounds = 0;
                    if (RESUME) {
                        uint32_t u_cur, u_sp, u_best, u_dpos;  // temporaries: T itself must never have its address taken
                        float u_end;
                        unpark_lane(pe, stk, blockDim.x, u_cur, u_sp, u_best, u_end, u_dpos);
                        T.cur = u_cur; T.sp = u_sp; T.best = u_best; T.end = u_end; T.dpos = u_dpos;
                    } else if (MODE == 3) {
                        // degenerate direction, or k_shade sent the segment back (fast-path hit not provable): reference tree only
                        if (!fast_eligible(T) || (__float_as_uint(b.w) & RT_FORCE_REF)) T.dpos |= RT_REFMODE_BIT;
                        else trav_setup_fast(S, T);
                    }
                }
            }
            batch_next += min(avail, (uint32_t)__popc(idle));
            act_mask = __ballot_sync(FULL, active);
        }
        if (act_mask == 0u) {
            if (exhausted && batch_next == batch_end) break;
            continue;
        }
        if (COUNT && MODE == 3) {
            const uint32_t rm = __ballot_sync(FULL, active && (T.dpos & RT_REFMODE_BIT));
            if (rm) { rrounds_c += 1u; rlanes_c += __popc(rm); }
        }
        if (COUNT && lane == 0) {
            rounds_c += 1u; busy_c += __popc(act_mask);
            if (exhausted && batch_next == batch_end) { trounds_c += 1u; tbusy_c += __popc(act_mask); }
        }
        // phase vote: run the leaf phase when (weighted) more lanes wait at a leaf than at an internal node
        const uint32_t n_leaf = __popc(__ballot_sync(FULL, active && (T.cur & RT_REF_LEAF)));
        const bool do_leaf = n_leaf * R.leaf_weight > ((uint32_t)__popc(act_mask) - n_leaf) * 16u;
        if (active) {
            bool more;
            if (COUNT) ++my_rounds;
            if constexpr (MODE == 3) {
                if (T.dpos & RT_REFMODE_BIT) more = trav_round<2, COUNT>(S, T, do_leaf, stk, blockDim.x, &tc, ExtendKey<FEAT>{&R, q, my_pos});
                else more = trav_round_fast<COUNT>(S, T, do_leaf, stk, blockDim.x, &tc, ExtendKey<FEAT>{&R, q, my_pos});
            } else {
                more = trav_round<MODE, COUNT>(S, T, do_leaf, stk, blockDim.x, &tc, ExtendKey<FEAT>{&R, q, my_pos});
            }
            if (!more) {
                if (MODE == 3 && !(T.dpos & RT_REFMODE_BIT) && T.best != RT_NO_HIT && (T.dpos & RT_TIE_BIT) && !R.debug_skip_redo) {
                    // an equal-t tie the fast path could not resolve (scenes with wrapped elements / volumes): same ray again on the
                    // reference tree, in place (the root test already passed in the producer)
                    T.best = RT_NO_HIT; T.end = INF; T.sp = 0; T.cur = S.root;
                    T.dpos = (T.d.x > 0.0f ? 1u : 0u) | (T.d.y > 0.0f ? 2u : 0u) | (T.d.z > 0.0f ? 4u : 0u) | RT_REFMODE_BIT;
                    if (COUNT) tc.redo++;
                } else {
                    // A fast-path hit still has to be PROVEN to be the reference's result (ref_proven).  That proof runs in k_shade, where
                    // all 32 lanes hold a hit, instead of here with the one or two lanes that finish in a given round (it was 5.5 % of
                    // this kernel's instructions at 1.2 active lanes): the hit is stored without its leaf bit = "unverified".
                    uint32_t out_ref = T.best;
                    if (MODE == 3 && !(T.dpos & RT_REFMODE_BIT) && T.best != RT_NO_HIT && !R.debug_skip_redo) out_ref &= ~RT_REF_LEAF;
                    R.hits[my_pos] = make_uint2(out_ref, __float_as_uint(T.end));
                    active = false;
                    if (COUNT) atomicAdd(&R.ctr->hist[min(23, 31 - __clz(max(my_rounds, 1u)))], 1ull);
                    if (COUNT && R.long_rays && my_rounds >= 4096u) {
                        const uint32_t slot = atomicAdd(reinterpret_cast<uint32_t*>(R.long_rays), 1u);
                        if (slot < 4095u) {
                            float* lr = R.long_rays + 8u * (slot + 1u);
                            lr[0] = T.o.x; lr[1] = T.o.y; lr[2] = T.o.z; lr[3] = T.d.x; lr[4] = T.d.y; lr[5] = T.d.z;
                            lr[6] = (float)my_rounds; lr[7] = (MODE == 3 && (T.dpos & RT_REFMODE_BIT)) ? 1.0f : 0.0f;
                        }
                    }
                }
            }
        }
        act_mask = __ballot_sync(FULL, active);
        // Leave with rays still running (they are parked below) once the warp has had its budget of rounds without being able to
        // refill.  The round count lives in batch_next == batch_end, which have no other use once nothing is left to claim.
        if (!RESUME) {
            if (R.tail_budget != 0u && exhausted && batch_next == batch_end && act_mask != 0u) {
                if (batch_end < RT_TAIL_BASE) batch_end = RT_TAIL_BASE;
                batch_next = ++batch_end;
                if ((batch_end & 63u) == 0u && batch_end - RT_TAIL_BASE > __ldcg(&R.ctr->budget)) break;
            }
        } else {
            // a resumed traversal runs while the queue launch beside it does, and at least `budget` rounds
            if (act_mask != 0u && ((++resume_rounds) & 63u) == 0u && resume_rounds > __ldcg(&R.ctr->budget) && __ldcg(&R.ctr->main_warps_left) == 0u) break;
        }
    }
    if (act_mask == 0u) break;
    {
        // park (RenderDev::susp): cursor, best hit and stack go to pool q^1; k_shade relinks the entry to the path's next position
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&R.ctr->n_susp[q ^ 1], (uint32_t)__popc(act_mask));
        base = __shfl_sync(FULL, base, 0);
        const uint32_t idx = base + __popc(act_mask & ((1u << lane) - 1u));
        if (active && idx < R.susp_cap) {
            park_lane(R.susp[q ^ 1] + (size_t)idx * R.susp_stride, stk, blockDim.x, my_pos, T.cur, T.sp, T.best, T.end, T.dpos);
            R.hits[my_pos] = make_uint2(RT_SUSPENDED, idx);
            active = false;
        }
        act_mask = __ballot_sync(FULL, active);
        if (act_mask == 0u) break;
        // pool full (1 Ki .. 64 Ki entries per pipeline): whoever could not be parked gets another budget
        if (RESUME) resume_rounds = 0u; else batch_next = batch_end = RT_TAIL_BASE + 1u;
    }
    }
    if (!RESUME && lane == 0) {
        atomicSub(&R.ctr->main_warps_left, 1u);  // the RESUME launch beside this one parks its rays when all are gone
        if (batch_end >= RT_TAIL_BASE) atomicAdd(&R.ctr->tail_acc, (1ull << 40) | (unsigned long long)(batch_end - RT_TAIL_BASE));  // -> next budget
    }
    if (COUNT) {
        for (int off = 16; off > 0; off >>= 1) {
            tc.nodes += __shfl_xor_sync(FULL, tc.nodes, off);
            tc.prims += __shfl_xor_sync(FULL, tc.prims, off);
            tc.exact += __shfl_xor_sync(FULL, tc.exact, off);
            tc.redo += __shfl_xor_sync(FULL, tc.redo, off);
            tc.max_sp = max(tc.max_sp, __shfl_xor_sync(FULL, tc.max_sp, off));
        }
        if (lane == 0) {
            atomicAdd(&R.ctr->node_visits, (unsigned long long)tc.nodes);
            atomicAdd(&R.ctr->prim_tests, (unsigned long long)tc.prims);
            atomicAdd(&R.ctr->exact_steps, (unsigned long long)tc.exact);
            atomicAdd(&R.ctr->redo_total, (unsigned long long)tc.redo);
            atomicMax(&R.ctr->max_sp, tc.max_sp);
            atomicAdd(&R.ctr->rounds, rounds_c); atomicAdd(&R.ctr->busy_lanes, busy_c);
            atomicAdd(&R.ctr->tail_rounds, trounds_c); atomicAdd(&R.ctr->tail_busy_lanes, tbusy_c);
            atomicAdd(&R.ctr->ref_rounds, rrounds_c); atomicAdd(&R.ctr->ref_lanes, rlanes_c);
        }
    }
}

// One bounce of ray_color (rendering.rs:25-68) for ONE path whose closest hit is `h`, including the accumulation of a terminated
// sample.  Returns true when the path continues (P then holds the scattered ray), false when the sample has ended.
struct PathIO {
    float3 o, d, thr;
    float dy0, time;
    uint32_t pixel, k, m;
    int32_t depth;
};
__device__ __forceinline__ bool shade_path(const SceneDev& S, const RenderDev& R, PathIO& P, const uint2 h) {
    float3 o = P.o, d = P.d, thr = P.thr;
    const float dy0 = P.dy0, time = P.time;
    const uint32_t k = P.k, m = P.m;
    int32_t depth = P.depth;
    const uint32_t pixel = P.pixel;  // the path carries its pixel (an integer modulo here was 4 % of the shade kernel's instructions and 13 % of its stalls)
    const uint32_t lane_id = R.lane_first + m * R.lane_stride;
    const float t = __uint_as_float(h.y);
    float3 radiance = f3(0.0f, 0.0f, 0.0f);
    bool terminated = true;
    // background_color.rs:9-19, sampled with the PRIMARY ray (Q6)
    float3 bgc;
    if (S.bg_is_sky) {
        float tt = 0.5f * (dy0 + 1.0f);
        float omt = 1.0f - tt;
        bgc = f3(1.0f * omt + 0.5f * tt, 1.0f * omt + 0.7f * tt, 1.0f * omt + 1.0f * tt);
    } else {
        bgc = f3(S.bg[0], S.bg[1], S.bg[2]);
    }
    if (R.mode == 1u) {  // RenderMode::Normals (rendering.rs:108-116)
        if (h.x != RT_NO_HIT) {
            Interaction hi = make_interaction(S, h.x, t, o, d, time, false);
            radiance = f3((hi.normal.x + 1.0f) * 0.5f, (hi.normal.y + 1.0f) * 0.5f, (hi.normal.z + 1.0f) * 0.5f);
        } else {
            radiance = bgc;
        }
    } else if (h.x == RT_NO_HIT) {
        radiance = f3(0.0f + thr.x * bgc.x, 0.0f + thr.y * bgc.y, 0.0f + thr.z * bgc.z);  // rendering.rs:66-68
    } else if (depth <= 1) {
        radiance = f3(0.0f, 0.0f, 0.0f);  // Q7, rendering.rs:26-27
    } else {
        Interaction hi = make_interaction(S, h.x, t, o, d, time, false);
        const MaterialDev m = S.materials[hi.material];
        const uint32_t s_global = k * R.n_tasks + lane_id;
        const uint4 rnd = philox4x32_10(R.seed0, R.seed1, pixel, s_global, 1u + (uint32_t)(R.max_depth - depth), 0u);
        float3 att = f3(1.0f, 1.0f, 1.0f), ndir = f3(0, 0, 1);
        float prob = 1.0f;
        bool scattered = false;
        float3 emitted = f3(0.0f, 0.0f, 0.0f);
        // texture coordinates are only needed by image textures: sphere uv costs acos+atan2
        float tu = hi.u, tv = hi.v;
        if (m.kind != 2u && ((h.x >> RT_REF_KIND_SHIFT) & 3u) == 0u && texture_needs_uv(S, m.texture)) {
            const Interaction hu = make_interaction(S, h.x, t, o, d, time, true);  // sphere uv costs acos + atan2: only on demand
            tu = hu.u; tv = hu.v;
        }
        if (m.kind == 0u) {  // Lambert, material.rs:59-63 + rendering.rs:35-51
            att = texture_sample(S, m.texture, tu, tv, hi.position);
            const float3 nrm = hi.normal;
            const float PI = 3.14159265358979323846f;
            float pdf;
            if (S.has_light) {  // sample_final_scattering_distribution, rendering.rs:73-86
                int p0, p1, pn;
                rect_axes(S.light_plane, p0, p1, pn);
                if (u01(rnd.x) < 0.5f) {  // gen_bool(0.5) -> RectGeometry::generate, rect_geometry.rs:60-67
                    float e0 = S.light_r0lo + u01(rnd.y) * (S.light_r0hi - S.light_r0lo);
                    float e1 = S.light_r1lo + u01(rnd.z) * (S.light_r1hi - S.light_r1lo);
                    float ex = p0 == 0 ? e0 : (p1 == 0 ? e1 : S.light_dist);
                    float ey = p0 == 1 ? e0 : (p1 == 1 ? e1 : S.light_dist);
                    float ez = p0 == 2 ? e0 : (p1 == 2 ? e1 : S.light_dist);
                    ndir = unit3(sub3(f3(ex, ey, ez), hi.position));
                } else {
                    float3 sph = sample_unit_sphere(u01(rnd.y), u01(rnd.z));
                    float3 v = add3(nrm, sph);
                    float l2 = dot3(v, v);
                    ndir = l2 > 1e-8f ? div3(v, sqrtf(l2)) : nrm;  // unit_or_else, vec3.rs:223-230
                }
                // RectGeometry::value, rect_geometry.rs:68-85
                float lv = 0.0f, lt;
                if (rect_t(S.light_plane, S.light_dist, S.light_r0lo, S.light_r0hi, S.light_r1lo, S.light_r1hi,
                           hi.position, ndir, 0.001f, __int_as_float(0x7f800000), lt)) {
                    float area = (S.light_r0hi - S.light_r0lo) * (S.light_r1hi - S.light_r1lo);
                    float cosine = fabsf(comp3(ndir, pn));
                    lv = (lt * lt) / (cosine * area);
                }
                pdf = 0.5f * lv + 0.5f * (fmaxf(dot3(nrm, ndir), 0.0f) / PI);
            } else {
                float3 sph = sample_unit_sphere(u01(rnd.x), u01(rnd.y));
                float3 v = add3(nrm, sph);
                float l2 = dot3(v, v);
                ndir = l2 > 1e-8f ? div3(v, sqrtf(l2)) : nrm;
                pdf = fmaxf(dot3(nrm, ndir), 0.0f) / PI;
            }
            float scattering_pdf = fmaxf(dot3(nrm, ndir), 0.0f) / PI;  // material.rs:123-127
            prob = scattering_pdf / pdf;                                // Q9: 0/0 -> NaN
            scattered = true;
        } else if (m.kind == 1u) {  // Metal, material.rs:65-79
            float3 fuzz_dir = f3(0.0f, 0.0f, 0.0f);
            if (m.param > 0.0f) {
                float3 sph = sample_unit_sphere(u01(rnd.x), u01(rnd.y));
                fuzz_dir = mul3(mul3(sph, cbrt_unit(u01(rnd.z))), m.param);
            }
            float3 refl = sub3(d, mul3(hi.normal, 2.0f * dot3(d, hi.normal)));  // reflect, vec3.rs:232-234
            float3 dir = add3(refl, fuzz_dir);
            if (dot3(dir, hi.normal) > 0.0f) {
                ndir = unit3(dir);
                att = texture_sample(S, m.texture, tu, tv, hi.position);
                scattered = true;
            }
        } else if (m.kind == 2u) {  // Dielectric, material.rs:80-105
            float ratio = hi.front_face ? 1.0f / m.param : m.param;
            float cos_theta = fminf(dot3(neg3(d), hi.normal), 1.0f);
            float sin_theta = sqrtf(1.0f - cos_theta * cos_theta);
            bool cannot_refract = ratio * sin_theta > 1.0f;
            bool do_reflect = cannot_refract;
            if (!do_reflect) {
                float r0 = (1.0f - ratio) / (1.0f + ratio);
                float rs = r0 * r0;
                float x = 1.0f - cos_theta;
                float x2 = x * x;
                float x4 = x2 * x2;
                float refl = rs + (1.0f - rs) * (x * x4);  // powi(5)
                do_reflect = refl > u01(rnd.x);
            }
            float3 dir;
            if (do_reflect) {
                dir = sub3(d, mul3(hi.normal, 2.0f * dot3(d, hi.normal)));
            } else {  // refract, vec3.rs:235-240
                float ct = fminf(dot3(neg3(d), hi.normal), 1.0f);
                float3 perp = mul3(add3(d, mul3(hi.normal, ct)), ratio);
                float3 par = mul3(hi.normal, -sqrtf(fabsf(1.0f - dot3(perp, perp))));
                dir = add3(perp, par);
            }
            ndir = unit3(dir);
            scattered = true;
        } else if (m.kind == 4u) {  // Isotropic (volume phase function), material.rs:106-111: Mirror(UnitSphere), not normalised
            ndir = sample_unit_sphere(u01(rnd.x), u01(rnd.y));
            att = texture_sample(S, m.texture, tu, tv, hi.position);
            scattered = true;
        } else {  // DiffuseLight: never scatters, emits from both faces (Q14), material.rs:132-137
            emitted = texture_sample(S, m.texture, tu, tv, hi.position);
        }
        if (scattered) {
            // rendering.rs:54-61.  accum_emitted stays 0 unless the throughput is non-finite, in which case
            // it turns NaN; folding that into the throughput (inf -> NaN) keeps every outcome identical.
            float3 nt = f3((thr.x * att.x) * prob, (thr.y * att.y) * prob, (thr.z * att.z) * prob);
            const float qnan = __int_as_float(0x7fc00000);
            if (isinf(thr.x)) nt.x = qnan;
            if (isinf(thr.y)) nt.y = qnan;
            if (isinf(thr.z)) nt.z = qnan;
            thr = nt;
            o = hi.position;
            d = ndir;
            depth -= 1;
            terminated = false;
        } else {
            radiance = f3(0.0f + thr.x * emitted.x, 0.0f + thr.y * emitted.y, 0.0f + thr.z * emitted.z);  // :62-64
        }
    }
    if (terminated) {
        float4* ap = R.acc + ((size_t)m * R.npix + pixel);
        float4 s = __ldcg(ap);  // sequential per-pixel sum in sample order (rendering.rs:172-178); streamed through L2, never reused from L1
        s.x += radiance.x; s.y += radiance.y; s.z += radiance.z;
        s.w += radiance.x * radiance.x + radiance.y * radiance.y + radiance.z * radiance.z;  // second moment (RMSE bound)
        __stcg(ap, s);
    }
    P.o = o; P.d = d; P.thr = thr; P.depth = depth;
    return !terminated;
}

// One bounce of ray_color for every live path of queue `q`; survivors (and regenerated paths) are compacted
// into queue q^1 with one warp-aggregated atomic per warp.
template <bool PROOF>  // PROOF: the extend launches run traversal 3, so hits may arrive unverified (scenes that keep the reference tree never do)
__global__ void __launch_bounds__(128, 6) k_shade(const __grid_constant__ SceneDev S, const __grid_constant__ RenderDev R, int q) {
    const uint32_t nb = R.ctr->nb[q], n = nb + R.ctr->nn[q];
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        R.ctr->cursor_extend = 0;
        R.ctr->n_susp[q] = 0;  // pool q has been consumed by the extend that just ran; the NEXT extend fills it again
        // The budget for the NEXT launch: 4 x the mean number of rounds a warp of this launch ran after it could no longer refill (at
        // least the caller's floor).  An ordinary tail is left alone — parking costs the path an iteration — and only the rays far
        // beyond it are parked: the distant-sphere and coplanar-fan rays that take 10^3 - 10^4 rounds.  A launch that consists of long
        // rays only (traversal 0: 0.4 N nodes per ray) has a long mean tail and parks nothing.  Safety net: while more than 0.4 % of
        // an iteration's rays still get parked, the budget doubles on top of that.
        const uint32_t parked = min(R.ctr->n_susp[q ^ 1], R.susp_cap);
        const unsigned long long acc = R.ctr->tail_acc;
        const uint32_t tail_warps = (uint32_t)(acc >> 40);
        const uint32_t mean_tail = tail_warps ? (uint32_t)((acc & ((1ull << 40) - 1ull)) / tail_warps) : 0u;
        uint32_t boost = R.ctr->boost;
        if ((unsigned long long)parked * 256ull > n) boost = min(boost + 1u, 16u);
        else if (boost) boost -= 1u;
        const uint32_t bud = R.tail_fixed ? R.tail_budget  // tests: park after exactly this many rounds, whatever the tails look like
                                          : (uint32_t)min((unsigned long long)max(R.tail_budget, 4u * mean_tail) << boost, 1ull << 24);
        R.ctr->budget = bud; R.ctr->boost = boost; R.ctr->tail_acc = 0ull;
        R.ctr->main_warps_left = R.extend_warps;
        if (R.iter_log) { R.iter_log[3u * 16384u + R.iter] = parked; R.iter_log[8u * 16384u + R.iter] = bud; }
        atomicAdd(&R.ctr->rays, (unsigned long long)n);
        if (R.iter_log) {  // [iter] live paths, [CAP + 2 iter] cumulative node visits (lo/hi), [3 CAP + 2 iter] cumulative prim tests
            R.iter_log[R.iter] = n;
            R.iter_log[16384u + R.iter] = (uint32_t)(R.ctr->node_visits >> 6);
            R.iter_log[32768u + R.iter] = (uint32_t)(R.ctr->prim_tests >> 6);
            R.iter_log[4u * 16384u + R.iter] = (uint32_t)(R.ctr->rounds >> 4);
            R.iter_log[5u * 16384u + R.iter] = (uint32_t)(R.ctr->busy_lanes >> 8);
            R.iter_log[6u * 16384u + R.iter] = (uint32_t)(R.ctr->tail_rounds >> 4);
            R.iter_log[7u * 16384u + R.iter] = (uint32_t)(R.ctr->tail_busy_lanes >> 8);
        }
    }
    uint32_t new_paths = 0;
    for (uint32_t base = warp_global * 32u; base < n; base += n_warps * 32u) {
        const uint32_t i = base + lane;
        const uint32_t ipos = queue_pos(i, nb, R.cap);
        bool live = false, fresh = false, ended = false, again = false, parked = false;
        uint32_t pool_entry = 0;
        PathIO P;
        P.o = f3(0, 0, 0); P.d = f3(0, 0, 1); P.thr = f3(1, 1, 1);
        P.dy0 = 0.0f; P.time = 0.0f; P.pixel = 0; P.k = 0; P.m = 0; P.depth = 0;
        if (i < n) {
            const float4 a = R.A[q][ipos], b = R.B[q][ipos], c = R.C[q][ipos];
            const int2 dm = R.D[q][ipos];
            const uint2 h = R.hits[ipos];
            P.depth = dm.x; P.m = (uint32_t)dm.y;
            P.time = R.Tm[q] ? R.Tm[q][ipos] : 0.0f;
            P.o = f3(a.x, a.y, a.z); P.d = f3(b.x, b.y, b.z); P.thr = f3(c.x, c.y, c.z);
            P.dy0 = c.w;
            P.pixel = __float_as_uint(a.w);
            P.k = __float_as_uint(b.w) & RT_K_MASK;
            uint2 hv = h;
            pool_entry = hv.y;
            if (hv.x == RT_SUSPENDED) {  // traversal parked by k_extend: the path waits, untouched, for the next extend to finish the segment
                again = true; parked = true;
                atomicAdd(&R.ctr->rays, ~0ull);
            } else if (PROOF && hv.x != RT_NO_HIT && !(hv.x & RT_REF_LEAF)) {  // unverified fast-path hit (k_extend): prove it here, with the whole warp at work
                hv.x |= RT_REF_LEAF;
                if (!ref_proven(S, hv.x, P.o, P.d, __uint_as_float(hv.y))) again = true;
            }
            if (!again) {
                live = shade_path(S, R, P, hv);
                ended = !live;
            } else if (parked) {
                live = true;
            } else {
                // not provable (~1e-5 of the rays): the path goes back into the queue unchanged and this segment is traced again, on
                // the reference tree; it is the same ray segment, so it is not counted twice
                live = true;
                atomicAdd(&R.ctr->rays, ~0ull);
                atomicAdd(&R.ctr->redo_total, 1ull);
            }
        }
        // A sample has ended: the slot renders the next sample of its chain, or, when the chain is complete, takes the next
        // chain of the pipeline (one warp-aggregated atomic) — the wavefront stays full while chains are left.
        const bool chain_done = ended && !(P.k + 1u < lane_samples(R, R.lane_first + P.m * R.lane_stride));
        const uint32_t mchain = __ballot_sync(0xFFFFFFFFu, chain_done);
        if (mchain) {
            unsigned long long cbase = 0;
            if (lane == 0) cbase = atomicAdd(&R.ctr->next_chain, (unsigned long long)__popc(mchain));
            cbase = __shfl_sync(0xFFFFFFFFu, cbase, 0);
            if (chain_done) {
                const unsigned long long cid = cbase + __popc(mchain & ((1u << lane) - 1u));
                if (cid < R.n_chains) {
                    uint32_t lane_index;
                    if (R.n_chains <= 0xFFFFFFFFull) {  // the usual case: 32-bit division
                        lane_index = (uint32_t)cid / R.npix;
                        P.pixel = (uint32_t)cid - lane_index * R.npix;
                    } else {
                        lane_index = (uint32_t)(cid / R.npix);
                        P.pixel = (uint32_t)(cid - (unsigned long long)lane_index * R.npix);
                    }
                    P.m = R.pipe + R.n_pipes * lane_index;
                    P.k = 0u;
                    fresh = true;
                }
            }
        }
        if (ended && !chain_done) { P.k += 1u; fresh = true; }
        if (fresh) {
            generate_path(S, R, P.pixel, R.lane_first + P.m * R.lane_stride, P.k, P.o, P.d, P.time);
            P.thr = f3(1.0f, 1.0f, 1.0f);
            P.dy0 = P.d.y;
            P.depth = R.max_depth;
            live = true;
            new_paths += 1u;
        }
        // warp-ballot compaction into the other queue: continuing paths from the bottom, fresh camera paths from the
        // top, so warps of the next extend hold either coherent primary rays or bounce rays, not a mix
        const uint32_t mcont = __ballot_sync(0xFFFFFFFFu, live && !fresh);
        const uint32_t mnew = __ballot_sync(0xFFFFFFFFu, live && fresh);
        if (mcont | mnew) {
            uint32_t bc = 0, bn = 0;
            if (lane == 0) {
                if (mcont) bc = atomicAdd(&R.ctr->nb[q ^ 1], (uint32_t)__popc(mcont));
                if (mnew) bn = atomicAdd(&R.ctr->nn[q ^ 1], (uint32_t)__popc(mnew));
            }
            bc = __shfl_sync(0xFFFFFFFFu, bc, 0);
            bn = __shfl_sync(0xFFFFFFFFu, bn, 0);
            if (live) {
                // a path that goes back for the same segment (parked / to be traced again) has passed the root test before
                const uint32_t miss = again ? (parked ? RT_RESUME : RT_FORCE_REF)
                                            : (root_gate(S, P.o, P.d, 0.001f, __int_as_float(0x7f800000), R.ref_only != 0u) ? 0u : RT_ROOT_MISS);
                const uint32_t below = (1u << lane) - 1u;
                const uint32_t pos = fresh ? R.cap - 1u - (bn + __popc(mnew & below)) : bc + __popc(mcont & below);
                R.A[q ^ 1][pos] = make_float4(P.o.x, P.o.y, P.o.z, __uint_as_float(P.pixel));
                R.B[q ^ 1][pos] = make_float4(P.d.x, P.d.y, P.d.z, __uint_as_float(P.k | miss));
                if (parked) R.susp[q ^ 1][(size_t)pool_entry * R.susp_stride].x = pos;  // link the parked traversal to the path's new position
                R.C[q ^ 1][pos] = make_float4(P.thr.x, P.thr.y, P.thr.z, P.dy0);
                R.D[q ^ 1][pos] = make_int2(P.depth, (int)P.m);
                if (R.Tm[q ^ 1]) R.Tm[q ^ 1][pos] = P.time;
            }
        }
    }
    // one atomic per warp for the generated-paths counter
    for (int off = 16; off > 0; off >>= 1) new_paths += __shfl_xor_sync(0xFFFFFFFFu, new_paths, off);
    if (lane == 0 && new_paths) atomicAdd(&R.ctr->paths, (unsigned long long)new_paths);
}


// film[p] = sum over this rank's lanes, in merge_planes order (last plane first), of lane_sum / lane_samples; times `scale`.
__global__ void __launch_bounds__(256) k_finalize(const __grid_constant__ RenderDev R, float* __restrict__ out_rgb, float scale,
                                                  int apply_scale, float* __restrict__ out_sumsq) {
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < R.npix; p += gridDim.x * blockDim.x) {
        float3 sum = f3(0, 0, 0);
        float sq = 0.0f;
        bool first = true;
        // order: lane T'-1 (if local), then ascending
        for (uint32_t step = 0; step <= R.n_local_lanes; ++step) {
            uint32_t local;
            if (step == 0) {
                uint32_t last_id = R.lane_first + (R.n_local_lanes - 1u) * R.lane_stride;
                if (last_id != R.n_tasks - 1u) continue;
                local = R.n_local_lanes - 1u;
            } else {
                local = step - 1u;
                uint32_t id = R.lane_first + local * R.lane_stride;
                if (id == R.n_tasks - 1u) continue;
            }
            uint32_t lane_id = R.lane_first + local * R.lane_stride;
            float ns = (float)lane_samples(R, lane_id);
            float4 a = R.acc[(size_t)local * R.npix + p];
            float3 plane = f3(a.x / ns, a.y / ns, a.z / ns);  // `.sum::<Color>() / real_samples as f32`
            if (first) { sum = plane; first = false; }
            else sum = add3(sum, plane);
            sq += a.w;
        }
        if (out_sumsq) out_sumsq[p] = sq;
        if (apply_scale) sum = mul3(sum, scale);
        out_rgb[3ull * p + 0] = sum.x;
        out_rgb[3ull * p + 1] = sum.y;
        out_rgb[3ull * p + 2] = sum.z;
    }
}

// Multi-GPU film combine fused with finalize (rt_render_multi): every rank holds the per-lane sums of ITS lanes (lane id % n_ranks
// == rank, local index id / n_ranks) in its own HBM; the GPU that owns a pixel slice reads all T' lane sums of those pixels —
// its own from HBM, the others through NVLink peer loads — and folds them in merge_planes order (last plane first, then
// ascending; rendering.rs:239-252) with the same arithmetic as k_finalize, so the film is bit-identical for ANY number of GPUs.
// The result is stored to every output pointer (own film, or all peers' films = the all-gather fused in as peer stores).
#define RT_MAX_GPUS 16
struct PeerPlanes {
    const float4* acc[RT_MAX_GPUS];  // acc[r]: rank r's lane sums, [lanes of r][npix]
    float* out[RT_MAX_GPUS];         // films to store the merged slice into (width*height*3 floats each)
    uint32_t n_ranks, n_out;
};
__global__ void __launch_bounds__(256) k_merge_peers(const __grid_constant__ PeerPlanes PP, const __grid_constant__ RenderDev R, uint32_t p_begin,
                                                     uint32_t p_end, float scale) {
    for (uint32_t p = p_begin + blockIdx.x * blockDim.x + threadIdx.x; p < p_end; p += gridDim.x * blockDim.x) {
        float3 sum = f3(0, 0, 0);
        for (uint32_t step = 0; step < R.n_tasks; ++step) {
            const uint32_t id = step == 0 ? R.n_tasks - 1u : step - 1u;  // lane T'-1 first, then 0 .. T'-2
            const float ns = (float)lane_samples(R, id);
            const float4 a = PP.acc[id % PP.n_ranks][(size_t)(id / PP.n_ranks) * R.npix + p];
            const float3 plane = f3(a.x / ns, a.y / ns, a.z / ns);
            sum = step == 0 ? plane : add3(sum, plane);
        }
        sum = mul3(sum, scale);
        for (uint32_t o = 0; o < PP.n_out; ++o) {
            float* out = PP.out[o];
            out[3ull * p + 0] = sum.x; out[3ull * p + 1] = sum.y; out[3ull * p + 2] = sum.z;
        }
    }
}

// Scene::hit on caller rays.
template <int MODE>
__global__ void __launch_bounds__(128) k_trace(const __grid_constant__ SceneDev S, const float* __restrict__ org,
                                               const float* __restrict__ dir, uint64_t n, float t_min, float t_max,
                                               uint32_t* __restrict__ hit_id, float* __restrict__ t_out,
                                               float* __restrict__ normal, uint8_t* __restrict__ front_face,
                                               float* __restrict__ uv) {
    extern __shared__ uint2 smem[];
    uint2* stk = reinterpret_cast<uint2*>(smem) + threadIdx.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        float3 o = f3(org[3 * i], org[3 * i + 1], org[3 * i + 2]);
        float3 d = f3(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]);
        HitRec h;
        bool proven = false;
        // Ray::new debug_asserts a unit direction (ray.rs:13) and SphereGeometry::hit relies on it (no d.d term): for a caller
        // ray outside that contract the reference arithmetic is not the geometric intersection, so no true-bound cull applies
        const float dd = dot3(d, d);
        if (MODE != 0 && !(fabsf(dd - 1.0f) <= 1e-5f)) {
            h = traverse<0, false>(S, o, d, S.trace_time, t_min, t_max, stk, blockDim.x, TraceKey{i});
        } else if constexpr (MODE == 3) {
            if (S.nodes2 && S.fast_trace_ok) h = traverse_fast(S, o, d, S.trace_time, t_min, t_max, stk, blockDim.x, &proven, TraceKey{i});
            if (!proven) h = traverse<2, false>(S, o, d, S.trace_time, t_min, t_max, stk, blockDim.x, TraceKey{i});
        } else {
            h = traverse<MODE, false>(S, o, d, S.trace_time, t_min, t_max, stk, blockDim.x, TraceKey{i});
        }
        if (h.ref != RT_NO_HIT) {
            Interaction hi = make_interaction(S, h.ref, h.t, o, d, S.trace_time, true);
            if (hit_id) hit_id[i] = h.ref & RT_REF_ITEM_MASK;
            if (t_out) t_out[i] = h.t;
            if (normal) { normal[3 * i] = hi.normal.x; normal[3 * i + 1] = hi.normal.y; normal[3 * i + 2] = hi.normal.z; }
            if (front_face) front_face[i] = hi.front_face ? 1 : 0;
            if (uv) { uv[2 * i] = hi.u; uv[2 * i + 1] = hi.v; }
        } else {
            if (hit_id) hit_id[i] = 0xFFFFFFFFu;
            if (t_out) t_out[i] = __int_as_float(0x7f800000);
            if (normal) { normal[3 * i] = 0.0f; normal[3 * i + 1] = 0.0f; normal[3 * i + 2] = 0.0f; }
            if (front_face) front_face[i] = 0;
            if (uv) { uv[2 * i] = 0.0f; uv[2 * i + 1] = 0.0f; }
        }
    }
}

__global__ void k_scale(float* __restrict__ x, uint64_t n, float m) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) x[i] = x[i] * m;
}

// Color::to_rgb8_gamma2 (color.rs:43-48): u8(clamp(0, 255, sqrt(c) * 256)), NaN -> 0
__global__ void k_gamma2(const float* __restrict__ rgb, uint64_t n, uint8_t* __restrict__ out) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        float v = sqrtf(rgb[i]) * 256.0f;
        float c = v > 255.0f ? 255.0f : (v < 0.0f ? 0.0f : v);  // math::clamp, math.rs:18-27 (NaN falls through)
        out[i] = (c == c) ? (uint8_t)c : (uint8_t)0;
    }
}

}  // namespace rt
