// rt_device.cuh — device data layout + device functions of the B200 render hot path.
//
// Everything in here is compiled with --fmad=false: the reference is Rust (never contracts a*b+c), so
// every float op below is a single IEEE-rounded op in the reference's operand order.  Citations are
// file:line in the reference repo (martingleich/RayTracingInAWeekend).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rt {

// ------------------------------------------------------------------------------------------------
// Device scene layout (HBM, all 16-byte aligned, fetched with 128-bit loads)
// ------------------------------------------------------------------------------------------------
// Child reference: bit31 = leaf, bits 30..29 = geometry kind, bit 28 = wrapped (run-time Transformation / Animation),
// bits 27..0 = item index (= hit id).
#define RT_REF_LEAF 0x80000000u
#define RT_REF_KIND_SHIFT 29
#define RT_REF_WRAPPED 0x10000000u
#define RT_REF_ITEM_MASK 0x0FFFFFFFu
#define RT_NO_HIT 0xFFFFFFFFu

// One internal BVH node = 64 B: the AABBs of BOTH children live in the parent ("children-in-parent"),
// so one node visit is one 64-byte fetch (4 x LDG.128) and two independent box tests.
//   q0 = Lmin.xyz Lmax.x | q1 = Lmax.yz Rmin.xy | q2 = Rmin.z Rmax.xyz | q3 = left, right, axis, guard band
//   guard band = two bf16 (A << 16 | C): m(t) = A + |t| * (1e-5 + C |t|), the max over both child subtrees
struct __align__(16) NodeDev {
    float4 q0, q1, q2;
    uint4 q3;
};
// One primitive = 64 B intersection record, indexed by item id.
//   sphere  : q0 = center.xyz, radius
//   rect    : q0 = dist, r0.lo, r0.hi, r1.lo | q1 = r1.hi, bits(plane), 0, 0
//   box     : q0 = min.xyz, 0 | q1 = max.xyz, 0
//   triangle: q0 = p0.xyz, n.x | q1 = n.yz, vt1.xy | q2 = vt1.z, d1, vt2.xy | q3 = vt2.z, d2, 0, 0
//             n = unit(cross(dir1,dir2)), vt1 = cross(n,dir2), d1 = dot(dir1,vt1), vt2 = cross(n,dir1), d2 = dot(dir2,vt2)
//             — the per-test constants of triangle_geometry.rs:14-29, hoisted to scene build with identical rounding.
struct __align__(16) PrimDev {
    float4 q0, q1, q2, q3;
};
// One shading record = 64 B, indexed by item id (read only for the final hit of a segment).
//   q0 = n0.xyz n1.x | q1 = n1.yz n2.xy | q2 = n2.z uv0.xy uv1.x | q3 = uv1.y uv2.xy bits(material)
struct __align__(16) ShadeDev {
    float4 q0, q1, q2, q3;
};
// Run-time wrappers of an item (row f2): SceneElement::Transformation{offset, y_sine, y_cosine} and ::Animation(velocity).
//   a = offset.xyz, y_sine | b = y_cosine, velocity.xyz | flags in the child reference (RT_REF_WRAPPED) + c.x
struct __align__(16) WrapDev {
    float4 a, b;
    uint32_t flags;          // RT_WRAP_* bits (1 Transformation, 2 Animation, 4 Volume)
    float neg_inv_density;   // Volume: -1 / density
    uint32_t pad1, pad2;
};
struct MaterialDev {  // 16 B
    uint32_t kind, texture;
    float param;
    uint32_t pad;
};
struct TextureDev {  // 64 B
    uint32_t kind, width, height, even;  // even / odd: RT_TEX_CHECKER children
    float r, g, b;
    uint32_t odd;
    const uint8_t* texels;  // RGB8, row 0 = top
    float inv_frequency;    // checker; marble: scale
    uint32_t pad;
    const float* ranvec;    // marble: 256 x 3 Perlin gradients
    const uint32_t* perm;   // marble: perm_x | perm_y | perm_z (256 each)
};

struct CameraDev {
    float3 position, ulc, unit_right, unit_up, scaled_right, scaled_up;
    float lens_radius;
    float time_start, time_end;
    float pace_x, pace_y;  // Camera::shutter_pace (rolling shutter, camera.rs:190); zero in every builder of the reference
};

struct SceneDev {
    const NodeDev* nodes;
    const PrimDev* prims;
    const ShadeDev* shade;
    const MaterialDev* materials;
    const TextureDev* textures;
    uint32_t root;            // child reference of the BVH root
    uint32_t n_items;
    float root_min[3], root_max[3];
    uint32_t stack_cap;       // traversal stack entries per thread
    uint32_t root_margin;     // guard band of the root box test (packed like NodeDev.q3.w)
    const float4* own_box;    // [2*n_nodes] each node's own AABB (min.xyz, bits(parent node | 0xFFFFFFFF for the root) | max.xyz, 0): read on
                              // exact-tie pops and by the ancestor walks (ref_reaches_exact, ref_first_visited)
    const WrapDev* wraps;     // [n_items] or NULL when no item is wrapped
    // Row f4 (rt_rebuild.hpp): the rebuilt topology over the same elements; NULL when the scene has wrapped elements
    const NodeDev* nodes2;    // same 64 B record; q3.z unused (children are visited near-first)
    uint32_t root2;
    const float4* vis_box;    // [2*n_items] box of the element's parent in the REFERENCE tree (min.xyz, bits(parent node) | max.xyz, 0)
    uint32_t fast_trace_ok;   // rt_trace may use the rebuilt tree (0 when the scene has Animation leaves: the tree's boxes only
                              // cover the camera's shutter interval, rt_trace takes any ray time)
    const uint32_t* visit_rank;  // [8][n_items] position of each element in the reference's ordered DFS for the 8 direction-sign octants
                              // (bit i = d_i > 0), or NULL: equal-t ties between elements are then resolved on the reference tree
    float trace_time;         // Ray::time of rt_trace rays
    CameraDev cam;
    uint32_t bg_is_sky;
    float bg[3];
    uint32_t has_light, light_plane;
    float light_dist, light_r0lo, light_r0hi, light_r1lo, light_r1hi;
};

// ------------------------------------------------------------------------------------------------
// small vector helpers (component order and association exactly as src/lib/vec3.rs)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
__device__ __forceinline__ float3 add3(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 sub3(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 neg3(float3 a) { return f3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float3 mul3(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float3 div3(float3 a, float s) { return f3(a.x / s, a.y / s, a.z / s); }
__device__ __forceinline__ float dot3(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }  // vec3.rs:189-191
__device__ __forceinline__ float len3(float3 a) { return sqrtf(dot3(a, a)); }
__device__ __forceinline__ float3 with_length3(float3 a, float len) { return mul3(a, len / len3(a)); }  // vec3.rs:206-208
__device__ __forceinline__ float3 unit3(float3 a) { return with_length3(a, 1.0f); }                      // vec3.rs:203-205
__device__ __forceinline__ float comp3(float3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

// 256-bit read-only global load (sm_100: LDG.E.256): a 64 B node / primitive record is two requests instead of four.  The
// extend kernel's deep-bounce launches are bound by L1TEX request throughput (ncu: 85 %), not by bytes.
__device__ __forceinline__ void ldg256(const void* p, float4& a, float4& b) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                 : "l"(p));
}
__device__ __forceinline__ void ldg_record(const void* p, float4& q0, float4& q1, float4& q2, float4& q3) {
    ldg256(p, q0, q1);
    ldg256(static_cast<const char*>(p) + 32, q2, q3);
}

// ------------------------------------------------------------------------------------------------
// Aabb::hit_cond (src/lib/aabb.rs:65-78) + carried-interval bounds
// ------------------------------------------------------------------------------------------------
struct Slab {
    bool ref_pass;  // the reference's per-axis-independent rule (Q1)
    float tn;       // max_i max(t0_i, start)   — also the exact end-dependent part of the rule: reject iff end <= tn
    float tf;       // min_i min(t1_i, end)
};
__device__ __forceinline__ void slab_axis(float mn, float mx, float o, float d, float& t0, float& t1) {
    float a = (mn - o) / d;  // true IEEE division, as the reference (no precomputed reciprocal)
    float b = (mx - o) / d;
    bool lt = a < b;  // math.rs:35-41 minmax: NaN compares false -> (b, a)
    t0 = lt ? a : b;
    t1 = lt ? b : a;
}
__device__ __forceinline__ Slab slab_test(float mnx, float mny, float mnz, float mxx, float mxy, float mxz, float3 o,
                                          float3 d, float start, float end) {
    float t0x, t1x, t0y, t1y, t0z, t1z;
    slab_axis(mnx, mxx, o.x, d.x, t0x, t1x);
    slab_axis(mny, mxy, o.y, d.y, t0y, t1y);
    slab_axis(mnz, mxz, o.z, d.z, t0z, t1z);
    float nx = fmaxf(t0x, start), ny = fmaxf(t0y, start), nz = fmaxf(t0z, start);  // f32::max ignores NaN, like fmaxf
    float fx = fminf(t1x, end), fy = fminf(t1y, end), fz = fminf(t1z, end);
    Slab s;
    s.ref_pass = !(fx <= nx) && !(fy <= ny) && !(fz <= nz);
    s.tn = fmaxf(fmaxf(nx, ny), nz);
    s.tf = fminf(fminf(fx, fy), fz);
    return s;
}

// ------------------------------------------------------------------------------------------------
// Primitive tests: return t (accepted iff start <= t < end, half-open like Range::contains) or NaN-free miss flag
// ------------------------------------------------------------------------------------------------
// SphereGeometry::hit, src/lib/hittable/sphere_geometry.rs:21-43 (t only; normal/uv are rebuilt for the final hit)
__device__ __forceinline__ bool sphere_t(float4 s, float3 o, float3 d, float start, float end, float& t_out) {
    float3 oc = sub3(o, f3(s.x, s.y, s.z));
    float half_b = dot3(oc, d);
    float c = dot3(oc, oc) - s.w * s.w;
    float disc = half_b * half_b - c;
    if (disc < 0.0f) return false;
    float sqrtd = sqrtf(disc);
    float root_small = -half_b - sqrtd;
    if (start <= root_small && root_small < end) { t_out = root_small; return true; }
    float root_large = -half_b + sqrtd;
    if (start <= root_large && root_large < end) { t_out = root_large; return true; }
    return false;
}
// RectGeometry::hit, src/lib/hittable/rect_geometry.rs:33-42
__device__ __forceinline__ void rect_axes(uint32_t plane, int& p0, int& p1, int& n) {
    if (plane == 0u) { p0 = 0; p1 = 1; n = 2; }
    else if (plane == 1u) { p0 = 0; p1 = 2; n = 1; }
    else { p0 = 1; p1 = 2; n = 0; }
}
__device__ __forceinline__ bool rect_t(uint32_t plane, float dist, float r0lo, float r0hi, float r1lo, float r1hi,
                                       float3 o, float3 d, float start, float end, float& t_out) {
    int p0, p1, n;
    rect_axes(plane, p0, p1, n);
    float t = (dist - comp3(o, n)) / comp3(d, n);
    if (start <= t && t < end) {
        float3 pos = add3(o, mul3(d, t));
        float a = comp3(pos, p0), b = comp3(pos, p1);
        if (a >= r0lo && a <= r0hi && b >= r1lo && b <= r1hi) { t_out = t; return true; }
    }
    return false;
}
// TriangleGeometry::hit, src/lib/hittable/triangle_geometry.rs:13-31 with the per-triangle constants precomputed
__device__ __forceinline__ bool triangle_t(const PrimDev& p, float3 o, float3 d, float start, float end, float& t_out,
                                           float& w1_out, float& w2_out) {
    float3 p0 = f3(p.q0.x, p.q0.y, p.q0.z);
    float3 n = f3(p.q0.w, p.q1.x, p.q1.y);
    float denom = dot3(d, n);
    if (fabsf(denom) > 0.0001f) {
        float t = dot3(sub3(p0, o), n) / denom;
        if (start <= t && t < end) {
            float3 pos = add3(o, mul3(d, t));
            float3 q = sub3(pos, p0);
            float3 vt1 = f3(p.q1.z, p.q1.w, p.q2.x);
            float w1 = dot3(q, vt1) / p.q2.y;
            if (w1 > 0.0f && w1 < 1.0f) {
                float3 vt2 = f3(p.q2.z, p.q2.w, p.q3.x);
                float w2 = dot3(q, vt2) / p.q3.y;
                float w0 = 1.0f - w1 - w2;
                if (w2 > 0.0f && w0 > 0.0f) { t_out = t; w1_out = w1; w2_out = w2; return true; }
            }
        }
    }
    return false;
}

__device__ __forceinline__ PrimDev load_prim(const PrimDev* p) {
    const float4* q = reinterpret_cast<const float4*>(p);
    PrimDev r;
    ldg_record(q, r.q0, r.q1, r.q2, r.q3);
    return r;
}

// Aabb::hit as a primitive (src/lib/aabb.rs:80-167): intersections_line + near/far choice.  plane_out = axis of the hit face.
__device__ __forceinline__ bool box_t(float4 mn4, float4 mx4, float3 o, float3 d, float start, float end, float& t_out, int& plane_out) {
    const float mn[3] = {mn4.x - o.x, mn4.y - o.y, mn4.z - o.z};
    const float mx[3] = {mx4.x - o.x, mx4.y - o.y, mx4.z - o.z};
    const float dd[3] = {d.x, d.y, d.z};
    float near = -__int_as_float(0x7f800000), far = __int_as_float(0x7f800000);
    int np = 0, fp = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float t1 = mn[i] / dd[i], t2 = mx[i] / dd[i];
        const float t_min = fminf(t1, t2), t_max = fmaxf(t1, t2);
        if (t_min > near) { near = t_min; np = i; }
        if (t_max < far) { far = t_max; fp = i; }
        if (near > far || far < 0.0f) return false;
    }
    if (start <= near && near < end) { t_out = near; plane_out = np; return true; }
    if (start <= far && far < end) { t_out = far; plane_out = fp; return true; }
    return false;
}
// rotate_around_up (transformations.rs:107-111)
__device__ __forceinline__ float3 rot_y(float c, float s, float3 v) { return f3(c * v.x + s * v.z, v.y, -s * v.x + c * v.z); }
// Ray seen by a wrapped element: Animation first (outer), then Transformation (hittable.rs:234-244, 270-277)
__device__ __forceinline__ void wrap_reverse_ray(const WrapDev& w, float time, float3& o, float3& d) {
    if (w.flags & 2u) {  // Animation: Transformation::ZERO.translate(velocity * time).reverse_ray
        const float3 off = f3(w.b.y * time, w.b.z * time, w.b.w * time);
        o = rot_y(1.0f, -0.0f, sub3(o, off));
        d = rot_y(1.0f, -0.0f, d);
    }
    if (w.flags & 1u) {
        o = rot_y(w.b.x, -w.a.w, sub3(o, f3(w.a.x, w.a.y, w.a.z)));
        d = rot_y(w.b.x, -w.a.w, d);
    }
}
// Transformation::apply_hit_interaction for both wrappers, inner first (hittable.rs:279-283)
__device__ __forceinline__ void wrap_apply_hit(const WrapDev& w, float time, float3& pos, float3& nrm) {
    if (w.flags & 1u) {
        pos = add3(rot_y(w.b.x, w.a.w, pos), f3(w.a.x, w.a.y, w.a.z));
        nrm = rot_y(w.b.x, w.a.w, nrm);
    }
    if (w.flags & 2u) {
        pos = add3(rot_y(1.0f, 0.0f, pos), f3(w.b.y * time, w.b.z * time, w.b.w * time));
        nrm = rot_y(1.0f, 0.0f, nrm);
    }
}
__device__ __forceinline__ WrapDev load_wrap(const SceneDev& S, uint32_t item) {
    const float4* q = reinterpret_cast<const float4*>(S.wraps + item);
    WrapDev w;
    w.a = __ldg(q); w.b = __ldg(q + 1);
    const float4 c = __ldg(q + 2);
    w.flags = __float_as_uint(c.x);
    w.neg_inv_density = c.y;
    w.pad1 = w.pad2 = 0u;
    return w;
}

__device__ __forceinline__ uint4 philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3);
__device__ __forceinline__ float u01(uint32_t v);
__device__ __forceinline__ float ln_unit(float u);

// Geometry::hit (hittable.rs:141-148): t of the accepted root in [start, end), by geometry kind.
__device__ __forceinline__ bool prim_t(uint32_t kind, const float4* q, float3 o, float3 d, float start, float end, float& t) {
    if (kind == 0u) return sphere_t(__ldg(q), o, d, start, end, t);
    if (kind == 2u) {
        PrimDev p;
        ldg_record(q, p.q0, p.q1, p.q2, p.q3);
        float w1, w2;
        return triangle_t(p, o, d, start, end, t, w1, w2);
    }
    if (kind == 3u) {
        int plane;
        return box_t(__ldg(q), __ldg(q + 1), o, d, start, end, t, plane);
    }
    const float4 a = __ldg(q), b = __ldg(q + 1);
    return rect_t(__float_as_uint(b.y), a.x, a.y, a.z, a.w, b.x, o, d, start, end, t);
}

// Key of the per-(path segment) Philox stream a volume draws its free path from (REPLAY layout of the oracle):
// philox(k0, k1; c0, c1, c2, 0x40000000 | item).x — keyed by the element, so it does not depend on the visit order.
struct VolKey {
    uint32_t k0, k1, c0, c1, c2;
};

// VolumeGeometry::hit (hittable.rs:315-342): two boundary tests, then t = max(start_medium, 0) - ln(U) / density.
template <class KEYFN>
__device__ __forceinline__ bool volume_t(uint32_t kind, const float4* q, uint32_t item, float neg_inv_density, float3 o, float3 d,
                                         float start, float end, const KEYFN& keyfn, float& t_out) {
    const float INF = __int_as_float(0x7f800000);
    float t0, t1;
    if (!prim_t(kind, q, o, d, -INF, INF, t0)) return false;
    const float next = t0 + 0.001f;
    if (!prim_t(kind, q, o, d, next, INF, t1)) return false;
    const float start_medium = fmaxf(t0, start), end_medium = fminf(t1, end);
    if (start_medium >= end_medium) return false;
    const VolKey k = keyfn();
    const uint4 r = philox4x32_10(k.k0, k.k1, k.c0, k.c1, k.c2, 0x40000000u | item);
    const float t = fmaxf(start_medium, 0.0f) + neg_inv_density * ln_unit(u01(r.x));
    if (t > end_medium) return false;
    t_out = t;
    return true;
}

// Leaf test dispatch (SceneElement::hit -> Geometry::hit, hittable.rs:141-148, 230-244).  Returns true and shrinks `end` when
// the element is accepted.  `time` is Ray::time (only Animation wrappers use it); keyfn supplies the volume draw key.
template <class KEYFN>
__device__ __forceinline__ bool leaf_test(const SceneDev& S, uint32_t ref, float3 o, float3 d, float time, float start, float& end,
                                          const KEYFN& keyfn) {
    const uint32_t kind = (ref >> RT_REF_KIND_SHIFT) & 3u;
    const uint32_t item = ref & RT_REF_ITEM_MASK;
    const float4* q = reinterpret_cast<const float4*>(S.prims + item);
    float t;
    bool hit;
    // KEYFN::FEAT: 0 = the scene has no wrapped element, 1 = Transformation / Animation wrappers, 2 = + volumes
    // (the launch picks the instantiation from the packed scene, so plain scenes do not carry the wrapper code)
    if constexpr (KEYFN::FEAT >= 1) {
        if (ref & RT_REF_WRAPPED) {
            const WrapDev w = load_wrap(S, item);
            wrap_reverse_ray(w, time, o, d);
            if constexpr (KEYFN::FEAT >= 2) {
                if (w.flags & 4u) hit = volume_t(kind, q, item, w.neg_inv_density, o, d, start, end, keyfn, t);
                else hit = prim_t(kind, q, o, d, start, end, t);
            } else {
                hit = prim_t(kind, q, o, d, start, end, t);
            }
        } else {
            hit = prim_t(kind, q, o, d, start, end, t);
        }
    } else {
        hit = prim_t(kind, q, o, d, start, end, t);
    }
    if (hit) end = t;
    return hit;
}

// ------------------------------------------------------------------------------------------------
// BoundingVolumeHierarchy::hit (src/lib/hittable.rs:429-473) as an iterative, ordered DFS, written as a
// per-lane state machine (trav_begin / trav_step) so a persistent warp can refill finished lanes.
//   * child order, the node test and the half-open acceptance are the reference's (Q13), so among equal-t
//     candidates the first visited wins exactly as in the recursion;
//   * the far child is pushed with tn; at pop the end-dependent part of hit_cond is re-evaluated
//     (reject iff end <= tn), which is what the recursion does when it reaches that child later;
//   * MODE 0: the reference rule alone, true divisions (the reference's algorithm, for A/B sweeps);
//     MODE 1: + tight cull: a child is skipped when the carried slab interval is empty by more than a per-node
//             guard band (true-bound AABBs cannot contain an accepted hit then), true divisions;
//     MODE 2: the same rules evaluated with one reciprocal per ray and multiplies, plus the per-axis padded
//             second-chance cull (padded_cull).  Whenever the approximate values cannot prove a reference-rule
//             decision (within RT_BAND relative, or the ray has a zero / denormal direction component) the exact
//             MODE 1 code runs for that node.  Every extra cull only removes children that cannot contain an accepted
//             hit, so the RESULT (element, t) is identical in all modes;
//     MODE 3: trav_round_fast below — a different tree, with a per-ray proof that the reference returns the same element.
// Stack: shared memory, [level][thread] entries of 8 B (child reference, cull threshold): one 64-bit access per push / pop.
// ------------------------------------------------------------------------------------------------
struct HitRec {
    uint32_t ref;  // child reference of the closest primitive, RT_NO_HIT if none
    float t;
};
struct TravCount {
    uint32_t nodes, prims, exact;  // node visits (64 B fetches), primitive tests, exact-fallback node steps
    uint32_t redo;                 // traversal 3: rays restarted on the reference tree
    uint32_t max_sp;               // deepest traversal stack use (entries); must stay <= SceneDev::stack_cap
};
struct Trav {
    float3 o, d;
    float time;        // Ray::time (Animation leaves)
    float3 inv, oinv;  // 1/d and -o/d: the fast path evaluates a slab plane as fma(b, inv, oinv)
    float start, end;
    float mscale;  // 2 * max_i |1 / d_i|: converts the lateral guard band into a gap in t
    float sB;      // 1e-5 * mscale
    float oe;      // 2 x bound on the absolute error of an approximate slab distance: 1.2e-7 * max_i |o_i / d_i| (x2)
    uint32_t cur, sp, best;
    uint32_t dpos;  // bit i = (d_i > 0): child order (hittable.rs:441)
    bool exact_only;
};
#define RT_BAND 1e-6f  // relative band of an approximate slab distance: 4x the bound 4 eps (see trav_setup)

// Guard band of the tight cull.  eps(t) = A + |t| * (1e-5 + C * |t|) is the LATERAL distance (space units) by which a ray
// may miss a primitive's true AABB and still be accepted by the primitive's single-precision test (A and C are packed
// as two bf16, rounded up: hi = A, lo = C).  A ray that misses a box laterally by delta on axis i opens a gap of
// delta / |d_i| in t, so the gap that proves "misses the box padded by eps" is eps * (1/|d_j| + 1/|d_k|) <= eps * mscale
// with mscale = 2 * max_i |1/d_i| (per ray).  Near-axis-parallel rays therefore (correctly) lose the tight cull:
// they can slide along a box face within rounding distance of it.
//
// Where to evaluate eps: an accepted hit at parameter t needs  t - g(t) <= tf  and  t + g(t) >= tn  with g(t) = eps(t) max_i|1/d_i|
// (every slab interval, padded, must contain t).  g grows with t, so the band must be taken at the LARGEST feasible t, which
// is near the interval's EXIT tf, not its entry tn: if g' <= 1/2 up to te = tf' + 2 g(tf') (tf' = max(tf, start)), then every
// feasible t is <= te, hence g(t) <= g(te) and the child can be skipped iff  tn - tf > 2 g(te).  (A ray nearly parallel to a
// slab has a huge tn for a box it misses; evaluating the quadratic sphere term there would make the band useless.)
// If g' > 1/2 at te (grazing rays, huge distances) the band is infinite: no tight cull.
// G2(t) := 2 g(t) = sA + t (sB + sC t)  with sA = A mscale (+ approximation error), sB = 1e-5 mscale, sC = C mscale.
// Independently of that, every accepted t is < end, so G2(end) is always a valid band once a hit has made end finite
// (band_at_end: the same for both children of a node, computed once).
__device__ __forceinline__ float band_at_end(float sA, float sB, float sC, float end) {
    return __fmaf_rn(end, __fmaf_rn(sC, end, sB), sA);  // +inf (or NaN from 0 * inf, ignored by fminf) while end = inf
}
__device__ __forceinline__ float guard_band(float sA, float sB, float sC, float tf, float start, float at_end) {
    const float tfp = fmaxf(tf, start);
    const float te = tfp + __fmaf_rn(tfp, __fmaf_rn(sC, tfp, sB), sA);
    const float mg = __fmaf_rn(te, __fmaf_rn(sC, te, sB), sA);
    return fminf((__fmaf_rn(2.0f * sC, te, sB) <= 1.0f) ? mg : __int_as_float(0x7f800000), at_end);  // NaN compares false: no cull
}
__device__ __forceinline__ float node_margin(uint32_t packed, float tf, float start, float end, float mscale) {
    const float a = __uint_as_float(packed & 0xFFFF0000u), c = __uint_as_float(packed << 16);
    return guard_band(a * mscale, 1e-5f * mscale, c * mscale, tf, start, band_at_end(a * mscale, 1e-5f * mscale, c * mscale, end));
}

// Root box test of BoundingVolumeHierarchy::hit (hittable.rs:472 -> hit_recursive(initial_index)): the reference rule, plus the
// tight cull unless ref_only.  Producers of rays (k_generate / k_shade, all 32 lanes busy) evaluate it once per ray so the
// persistent extend warps do not run 6 IEEE divisions with one or two lanes enabled.
__device__ __forceinline__ bool root_gate(const SceneDev& S, float3 o, float3 d, float start, float end, bool ref_only) {
    if (S.root & RT_REF_LEAF) return true;
    const Slab r = slab_test(S.root_min[0], S.root_min[1], S.root_min[2], S.root_max[0], S.root_max[1], S.root_max[2], o, d, start, end);
    bool pass = r.ref_pass;
    if (!ref_only) {
        const float mscale = 2.0f * fmaxf(fmaxf(fabsf(1.0f / d.x), fabsf(1.0f / d.y)), fabsf(1.0f / d.z));
        pass = pass && !(r.tn - r.tf > node_margin(S.root_margin, r.tf, start, end, mscale));
    }
    return pass;
}

// Per-ray traversal setup (the root box has already been tested: root_gate).
__device__ __forceinline__ void trav_setup(const SceneDev& S, Trav& T, float3 o, float3 d, float time, float start, float end) {
    T.o = o; T.d = d; T.time = time; T.start = start; T.end = end;
    T.best = RT_NO_HIT; T.sp = 0; T.cur = S.root;
    T.inv = f3(__frcp_rn(d.x), __frcp_rn(d.y), __frcp_rn(d.z));  // correctly rounded 1/d
    T.oinv = f3(-o.x * T.inv.x, -o.y * T.inv.y, -o.z * T.inv.z);
    T.mscale = 2.0f * fmaxf(fmaxf(fabsf(T.inv.x), fabsf(T.inv.y)), fabsf(T.inv.z));  // +inf for axis-parallel rays: no tight cull
    T.sB = 1e-5f * T.mscale;
    // fma(b, inv, oinv) vs the reference's rn(rn(b - o) / d): |err| <= 4 eps |t| + eps |o_i / d_i| = 2.4e-7 |t| + 6e-8 |o_i / d_i|
    // (inv, the fma and the reference's own two roundings act on t; the rounding of oinv acts on o / d).  The bands used are
    // RT_BAND |t| = 1e-6 |t| and T.oe = 2.4e-7 max_i |o_i / d_i|, i.e. 4x each (tests/test_slab_error_bound.py).
    T.oe = 2.4e-7f * fmaxf(fmaxf(fabsf(T.oinv.x), fabsf(T.oinv.y)), fabsf(T.oinv.z));
    T.dpos = (d.x > 0.0f ? 1u : 0u) | (d.y > 0.0f ? 2u : 0u) | (d.z > 0.0f ? 4u : 0u);
    const float lo = 1e-18f, hi = 1e18f;  // NaN / inf / zero / denormal components, or a negative range start -> exact path only
    T.exact_only = !(fabsf(d.x) > lo && fabsf(d.x) < hi && fabsf(d.y) > lo && fabsf(d.y) < hi && fabsf(d.z) > lo && fabsf(d.z) < hi) ||
                   !(start >= 0.0f) || !(T.oe < 1e30f);
}

template <int MODE>
__device__ __forceinline__ bool trav_begin(const SceneDev& S, Trav& T, float3 o, float3 d, float time, float start, float end) {
    trav_setup(S, T, o, d, time, start, end);
    return root_gate(S, o, d, start, end, MODE == 0);
}

// Approximate carried slab interval: 6 FFMA + 12 FMNMX.  FMA is deliberate here (this value only steers the cull /
// sure-pass decision, never a result); its error bound is T.oe + 1.2e-7 |t|.
__device__ __forceinline__ void approx_slab(float mnx, float mny, float mnz, float mxx, float mxy, float mxz, const Trav& T,
                                            float& tn, float& tf) {
    float ax = __fmaf_rn(mnx, T.inv.x, T.oinv.x), bx = __fmaf_rn(mxx, T.inv.x, T.oinv.x);
    float ay = __fmaf_rn(mny, T.inv.y, T.oinv.y), by = __fmaf_rn(mxy, T.inv.y, T.oinv.y);
    float az = __fmaf_rn(mnz, T.inv.z, T.oinv.z), bz = __fmaf_rn(mxz, T.inv.z, T.oinv.z);
    tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), T.start));
    tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), T.end));
}

// Second-chance tight cull, decisive for rays that run nearly parallel to an axis.  For them the scalar band
// eps * 2 max|1/d_i| is useless, and the reference's own rule is per-axis independent (Q1): it never intersects the slab
// intervals, so without a hit yet (end = inf) such a ray walks most of the tree — a handful of these rays set the duration
// of a whole launch.  Here every axis is padded with ITS OWN eps |1/d_i|: an accepted hit at t satisfies
// lo_i - eps(t)|1/d_i| <= t <= hi_i + eps(t)|1/d_i| on every axis (the hit point is within eps(t) of the element's true box),
// and t <= te := min(end, min_i(hi_i' + 2 g_i(hi_i'))) over the axes where g_i = eps |1/d_i| has slope <= 1/2.
// Returns true when the padded intervals do not intersect; tn_pad = padded entry (a valid pop threshold: t >= tn_pad).
__device__ __forceinline__ float padded_exit_bound(float h, float i, float a, float c, float start) {
    const float hp = fmaxf(h, start);
    const float te = __fmaf_rn(2.0f * i, __fmaf_rn(hp, __fmaf_rn(c, hp, 1e-5f), a), hp);
    return (i * __fmaf_rn(2.0f * c, te, 1e-5f) <= 0.5f) ? te : __int_as_float(0x7f800000);
}
struct PadRay {  // the ray fields padded_cull needs, by value (the function is deliberately not inlined: it is the rare path)
    float3 inv, oinv;
    float start, end;
};
__device__ __noinline__ float padded_cull_impl(float mnx, float mny, float mnz, float mxx, float mxy, float mxz, PadRay T, uint32_t packed) {
    const float a = __uint_as_float(packed & 0xFFFF0000u), c = __uint_as_float(packed << 16);
    const float ax = __fmaf_rn(mnx, T.inv.x, T.oinv.x), bx = __fmaf_rn(mxx, T.inv.x, T.oinv.x);
    const float ay = __fmaf_rn(mny, T.inv.y, T.oinv.y), by = __fmaf_rn(mxy, T.inv.y, T.oinv.y);
    const float az = __fmaf_rn(mnz, T.inv.z, T.oinv.z), bz = __fmaf_rn(mxz, T.inv.z, T.oinv.z);
    const float lx = fminf(ax, bx), hx = fmaxf(ax, bx), ly = fminf(ay, by), hy = fmaxf(ay, by), lz = fminf(az, bz), hz = fmaxf(az, bz);
    const float ix = fabsf(T.inv.x), iy = fabsf(T.inv.y), iz = fabsf(T.inv.z);
    float te = T.end;
    te = fminf(te, padded_exit_bound(hx, ix, a, c, T.start));
    te = fminf(te, padded_exit_bound(hy, iy, a, c, T.start));
    te = fminf(te, padded_exit_bound(hz, iz, a, c, T.start));
    const float eps_e = __fmaf_rn(te, __fmaf_rn(c, te, 1e-5f), a);  // inf / NaN when te = inf: fmaxf / fminf then drop the padded terms
    // error of an approximate slab distance ON ITS OWN AXIS: absolute part 2.4e-7 |o_i / d_i| (4x the bound eps |o_i / d_i|, trav_setup)
    // + 1e-6 relative (>= 8 ulp).  The scalar T.oe (the max over the axes) would pad every axis with the error of the worst one: for a
    // ray nearly parallel to an axis (|d_y| ~ 1e-8: |o_y / d_y| ~ 1e8, T.oe ~ 50) the x and z intervals would grow by +-50 in t and
    // the ray would walk every box that straddles its plane — 30 000 - 47 000 rounds, a whole launch waiting for one ray.
    const float oex = 2.4e-7f * fabsf(T.oinv.x), oey = 2.4e-7f * fabsf(T.oinv.y), oez = 2.4e-7f * fabsf(T.oinv.z);
    const float px = __fmaf_rn(eps_e, ix, oex), py = __fmaf_rn(eps_e, iy, oey), pz = __fmaf_rn(eps_e, iz, oez);
    const float nx = lx - __fmaf_rn(RT_BAND, fabsf(lx), px), ny = ly - __fmaf_rn(RT_BAND, fabsf(ly), py), nz = lz - __fmaf_rn(RT_BAND, fabsf(lz), pz);
    const float fx = hx + __fmaf_rn(RT_BAND, fabsf(hx), px), fy = hy + __fmaf_rn(RT_BAND, fabsf(hy), py), fz = hz + __fmaf_rn(RT_BAND, fabsf(hz), pz);
    const float tn_pad = fmaxf(fmaxf(nx, ny), fmaxf(nz, T.start));
    const float tf_pad = fminf(fminf(fx, fy), fminf(fz, T.end));
    return tn_pad > tf_pad ? __int_as_float(0x7f800000) : tn_pad;  // +inf = culled, else the padded entry
}
__device__ __forceinline__ bool padded_cull(float mnx, float mny, float mnz, float mxx, float mxy, float mxz, const Trav& T,
                                            uint32_t packed, float& tn_pad) {
    PadRay r;
    r.inv = T.inv; r.oinv = T.oinv; r.start = T.start; r.end = T.end;
    tn_pad = padded_cull_impl(mnx, mny, mnz, mxx, mxy, mxz, r, packed);
    return tn_pad == __int_as_float(0x7f800000);
}

// One scheduling round of a lane: a node step (if `do_leaf` is false and the lane is at an internal node) or a leaf
// step (if `do_leaf` is true and the lane is at a leaf), followed by the pops it triggers; lanes in the other state
// simply wait.  The warp picks `do_leaf` by vote (k_extend), so each phase runs with most lanes enabled instead of
// interleaving 17-lane node steps with 4-lane leaf tests.  Per-lane visit order is untouched.
// Returns false when the lane's traversal is finished.
template <int MODE, bool COUNT, class KEYFN>
__device__ __forceinline__ bool trav_round(const SceneDev& S, Trav& T, bool do_leaf, uint2* stk,
                                           uint32_t stk_stride, TravCount* cnt, const KEYFN& keyfn) {
    uint32_t cur = T.cur;
    const bool at_leaf = (cur & RT_REF_LEAF) != 0u;
    bool need_pop = false;
    if (do_leaf) {
        if (at_leaf) {
            if (COUNT) cnt->prims++;
            if (leaf_test(S, cur, T.o, T.d, T.time, T.start, T.end, keyfn)) T.best = cur;
            need_pop = true;
        }
    } else if (!at_leaf) {
        if (COUNT) cnt->nodes++;
        const float4* q = reinterpret_cast<const float4*>(S.nodes + cur);
        float4 q0, q1, q2, q3f;
        ldg_record(q, q0, q1, q2, q3f);
        const uint4 q3 = make_uint4(__float_as_uint(q3f.x), __float_as_uint(q3f.y), __float_as_uint(q3f.z), __float_as_uint(q3f.w));
        const bool leafL = (q3.x & RT_REF_LEAF) != 0u, leafR = (q3.y & RT_REF_LEAF) != 0u;
        float tnL, tfL, tnR, tfR, mgL = 0.0f, mgR = 0.0f;
        bool pL, pR, need_exact;
        if (MODE == 2 && !T.exact_only) {
            approx_slab(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, T, tnL, tfL);
            approx_slab(q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, T, tnR, tfR);
            // guard band in t (tn >= start >= 0 here): (A + 1e-5 t + C t^2) * mscale, plus the approximation error
            const float sA = __fmaf_rn(__uint_as_float(q3.w & 0xFFFF0000u), T.mscale, T.oe), sC = __uint_as_float(q3.w << 16) * T.mscale;
            const float at_end = band_at_end(sA, T.sB, sC, T.end);
            mgL = guard_band(sA, T.sB, sC, tfL, T.start, at_end);
            mgR = guard_band(sA, T.sB, sC, tfR, T.start, at_end);
            bool cullL = tnL - tfL > mgL, cullR = tnR - tfR > mgR;
            // second chance for a child whose carried interval is empty, but by less than the scalar band: the per-axis padded
            // test (strictly tighter; rare for ordinary rays, the normal case for rays nearly parallel to an axis)
            if (!cullL && tnL > tfL) {
                float pn;
                cullL = padded_cull(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, T, q3.w, pn);
                mgL = tnL - pn;  // a leaf entry is pushed with threshold tn - mg = the padded entry
            }
            if (!cullR && tnR > tfR) {
                float pn;
                cullR = padded_cull(q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, T, q3.w, pn);
                mgR = tnR - pn;
            }
            // carried interval non-empty beyond rounding => the per-axis reference rule passes too
            const bool sureL = leafL || (tfL - tnL > __fmaf_rn(RT_BAND, tfL + tnL, T.oe));
            const bool sureR = leafR || (tfR - tnR > __fmaf_rn(RT_BAND, tfR + tnR, T.oe));
            pL = !cullL && sureL;
            pR = !cullR && sureR;
            need_exact = (!cullL && !sureL) || (!cullR && !sureR);
        } else {
            need_exact = true;
        }
        if (need_exact) {
            if (COUNT) cnt->exact++;
            const Slab sl = slab_test(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, T.o, T.d, T.start, T.end);
            const Slab sr = slab_test(q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, T.o, T.d, T.start, T.end);
            // leaves are never box-tested by the reference (hittable.rs:436-437); internal children get hit_cond
            pL = leafL ? true : sl.ref_pass;
            pR = leafR ? true : sr.ref_pass;
            if (MODE >= 1) {
                pL = pL && !(sl.tn - sl.tf > node_margin(q3.w, sl.tf, T.start, T.end, T.mscale));
                pR = pR && !(sr.tn - sr.tf > node_margin(q3.w, sr.tf, T.start, T.end, T.mscale));
            }
            tnL = sl.tn; tnR = sr.tn;
            if (MODE >= 1) { mgL = node_margin(q3.w, sl.tf, T.start, T.end, T.mscale); mgR = node_margin(q3.w, sr.tf, T.start, T.end, T.mscale); }
        }
        const bool left_first = ((T.dpos >> q3.z) & 1u) != 0u;  // dir[axis_id] > 0, hittable.rs:441-445
        const uint32_t c1 = left_first ? q3.x : q3.y, c2 = left_first ? q3.y : q3.x;
        const bool p1 = left_first ? pL : pR, p2 = left_first ? pR : pL;
        if (p1 && p2) {
            float tn2 = left_first ? tnR : tnL;
            // leaf entries carry their cull threshold (tn - guard band), internal entries carry tn
            if (c2 & RT_REF_LEAF) tn2 = (MODE >= 1) ? tn2 - (left_first ? mgR : mgL) : -__int_as_float(0x7f800000);
            stk[T.sp * stk_stride] = make_uint2(c2, __float_as_uint(tn2));
            ++T.sp;
            if (COUNT) cnt->max_sp = max(cnt->max_sp, T.sp);
        }
        if (p1) T.cur = c1;
        else if (p2) T.cur = c2;
        else need_pop = true;
    }
    if (need_pop) {
        for (;;) {
            if (T.sp == 0) return false;
            --T.sp;
            const uint2 e = stk[T.sp * stk_stride];
            cur = e.x;
            const float tn = __uint_as_float(e.y);
            if (cur & RT_REF_LEAF) {
                if (T.end < tn) continue;  // tight cull with the shrunken range (never taken in MODE 0: tn = -inf)
                break;
            }
            // hit_cond's end-dependent part with the shrunken range: reject iff end <= tn (exact tn)
            if (MODE == 2 && !T.exact_only) {
                const float e = __fmaf_rn(RT_BAND, fabsf(tn), T.oe);  // relative + absolute part of the approximation error (trav_setup)
                if (T.end > tn + e) break;
                if (T.end < tn - e) continue;
                const float4 b0 = __ldg(S.own_box + 2u * cur), b1 = __ldg(S.own_box + 2u * cur + 1u);
                const Slab s = slab_test(b0.x, b0.y, b0.z, b1.x, b1.y, b1.z, T.o, T.d, T.start, T.end);
                if (T.end <= s.tn) continue;
                break;
            }
            if (T.end <= tn) continue;
            break;
        }
        T.cur = cur;
    }
    return true;
}

// ------------------------------------------------------------------------------------------------
// Row f4: traversal of the REBUILT topology (S.nodes2, plain scenes only).  It is an accelerator, not a definition:
//   * it finds the smallest accepted t over ALL elements (every cull is conservative: a child is skipped only when the
//     carried interval is empty by more than the guard band + the approximation error, and a popped entry only when
//     end < tn - band, so even an element at exactly the current best t is still tested);
//   * the leaf test runs with a CLOSED upper end once a best exists, so a second element at exactly the same t is seen
//     and flagged (`tie`): which of two equal-t elements the reference returns depends on ITS visit order (Q13);
//   * children are visited near-first (the reference orders by the sign of dir[axis_id]; irrelevant here).
// The caller accepts the result only if !tie and ref_visible(); everything else is traced again on the reference tree.
// ------------------------------------------------------------------------------------------------
#define RT_TIE_BIT 8u  // kept in Trav::dpos next to the direction signs (bits 0..2)

// Scenes with wrapped elements / volumes (KEYFN::FEAT >= 1): the rebuilt tree holds the TRUE world-space bounds of a wrapped
// element (the reference's own box for it is not a bound, Q2), the leaf test applies the wrappers exactly like leaf_test, and a
// volume's free-flight draw is keyed by the element, so it does not depend on which tree reaches it.  A volume accepts t <= end
// (hittable.rs:330: `t > end_medium` rejects), i.e. its range is closed by nature: an equal-t volume is flagged as a tie like
// everything else, and the reference tree decides.
template <bool COUNT, class KEYFN>
__device__ __forceinline__ bool trav_round_fast(const SceneDev& S, Trav& T, bool do_leaf, uint2* stk,
                                                uint32_t stk_stride, TravCount* cnt, const KEYFN& keyfn) {
    uint32_t cur = T.cur;
    const bool at_leaf = (cur & RT_REF_LEAF) != 0u;
    bool need_pop = false;
    if (do_leaf) {
        if (at_leaf) {
            if (COUNT) cnt->prims++;
            const uint32_t kind = (cur >> RT_REF_KIND_SHIFT) & 3u, item = cur & RT_REF_ITEM_MASK;
            const bool have = T.best != RT_NO_HIT;
            const float end_closed = have ? __uint_as_float(__float_as_uint(T.end + 0.0f) + 1u) : T.end;  // next float up (end >= 0, finite)
            const float4* q = reinterpret_cast<const float4*>(S.prims + item);
            float t;
            bool hit;
            bool done = false;
            float3 o = T.o, d = T.d;
            if constexpr (KEYFN::FEAT >= 1) {
                if (cur & RT_REF_WRAPPED) {
                    const WrapDev w = load_wrap(S, item);
                    wrap_reverse_ray(w, T.time, o, d);
                    if constexpr (KEYFN::FEAT >= 2) {
                        if (w.flags & 4u) {
                            hit = volume_t(kind, q, item, w.neg_inv_density, o, d, T.start, T.end, keyfn, t);
                            done = true;
                        }
                    }
                }
            }
            if (!done) hit = prim_t(kind, q, o, d, T.start, end_closed, t);
            if (hit) {
                if (have && t == T.end) {
                    // the same element can sit in several leaves (pre-split references): seeing it again is not a tie
                    if (((cur ^ T.best) & RT_REF_ITEM_MASK) != 0u) {
                        // Two elements accept exactly the same t (coplanar overlapping triangles: the Q4 fan of an OBJ mesh puts
                        // hundreds of them into the mesh's symmetry plane).  The reference keeps the one it visits FIRST (the later
                        // one fails t < end), and its visit order for this ray is fixed by the tree and the signs of the direction
                        // alone: every node visits the side dir[axis_id] > 0 picks first (hittable.rs:441-445).  S.visit_rank holds
                        // each element's position in that order for the 8 sign octants, so plain scenes resolve the tie right here
                        // with two loads; with wrapped elements / volumes (a volume accepts t <= end, so a LATER equal-t volume
                        // wins) the ray is flagged and traced again on the reference tree.
                        if (KEYFN::FEAT == 0 && S.visit_rank) {
                            const uint32_t* rk = S.visit_rank + (size_t)(T.dpos & 7u) * S.n_items;
                            if (__ldg(rk + (cur & RT_REF_ITEM_MASK)) < __ldg(rk + (T.best & RT_REF_ITEM_MASK))) T.best = cur;
                        } else {
                            T.dpos |= RT_TIE_BIT;
                        }
                    }
                } else {
                    T.best = cur; T.end = t; T.dpos &= ~RT_TIE_BIT;
                }
            }
            need_pop = true;
        }
    } else if (!at_leaf) {
        if (COUNT) cnt->nodes++;
        const float4* q = reinterpret_cast<const float4*>(S.nodes2 + cur);
        float4 q0, q1, q2, q3f;
        ldg_record(q, q0, q1, q2, q3f);
        const uint4 q3 = make_uint4(__float_as_uint(q3f.x), __float_as_uint(q3f.y), __float_as_uint(q3f.z), __float_as_uint(q3f.w));
        float tnL, tfL, tnR, tfR;
        approx_slab(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, T, tnL, tfL);
        approx_slab(q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, T, tnR, tfR);
        const float sA = __fmaf_rn(__uint_as_float(q3.w & 0xFFFF0000u), T.mscale, T.oe), sC = __uint_as_float(q3.w << 16) * T.mscale;
        // one band for both children, taken at the later of the two exits (the band grows with t: valid for both, 8 instructions less)
        const float mgB = guard_band(sA, T.sB, sC, fmaxf(tfL, tfR), T.start, band_at_end(sA, T.sB, sC, T.end));
        float mgL = mgB, mgR = mgB;
        bool pL = !(tnL - tfL > mgL), pR = !(tnR - tfR > mgR);
        if (pL && tnL > tfL) {  // empty by less than the scalar band: per-axis padded second chance (see padded_cull)
            float pn;
            pL = !padded_cull(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, T, q3.w, pn);
            mgL = tnL - pn;
        }
        if (pR && tnR > tfR) {
            float pn;
            pR = !padded_cull(q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, T, q3.w, pn);
            mgR = tnR - pn;
        }
        const bool left_first = tnL <= tnR;  // near-first (ordering by the split-axis sign, or as a tie-break, measured within 1 %)
        const uint32_t c1 = left_first ? q3.x : q3.y, c2 = left_first ? q3.y : q3.x;
        const bool p1 = left_first ? pL : pR, p2 = left_first ? pR : pL;
        if (p1 && p2) {
            stk[T.sp * stk_stride] = make_uint2(c2, __float_as_uint(left_first ? tnR - mgR : tnL - mgL));  // ref + cull threshold of the deferred child
            ++T.sp;
            if (COUNT) cnt->max_sp = max(cnt->max_sp, T.sp);
        }
        if (p1) T.cur = c1;
        else if (p2) T.cur = c2;
        else need_pop = true;
    }
    if (need_pop) {
        for (;;) {
            if (T.sp == 0) return false;
            --T.sp;
            const uint2 e = stk[T.sp * stk_stride];
            cur = e.x;
            if (T.end < __uint_as_float(e.y)) continue;
            break;
        }
        T.cur = cur;
    }
    return true;
}

// Can the fast traversal be used for this ray at all?  (finite, non-degenerate direction; see trav_setup)
// Can the fast traversal be used for this ray?  Not for degenerate directions (zero / denormal / non-finite components,
// trav_setup): the approximate slab arithmetic is not defined for them; the reference tree's exact path is.
__device__ __forceinline__ bool fast_eligible(const Trav& T) { return !T.exact_only; }
__device__ __forceinline__ void trav_setup_fast(const SceneDev& S, Trav& T) {
    T.cur = S.root2;
    T.dpos &= 7u;
}

// Does the reference traversal provably return element `ref` at `t` (given that it is the UNIQUE minimum of all accepted
// candidates)?  The reference reaches the element iff every ancestor A passes Aabb::hit_cond for the range [start, e) with
// some e > t.  Per axis that holds when t lies strictly between the two slab distances the reference computes,
// t0 = rn(rn(X - o) / d) for the entry face X (and t1 likewise).  With p = rn(o + rn(t d)) (four roundings of 2^-24 in
// total between the two sides) a sufficient condition is  p - X > 3.2 * 2^-24 (|X| + |o| + |p|)  = 1.9e-7 (...); we ask for
// 4e-7 on both faces of every axis.  Ancestor boxes are nested (exact unions) and the margin grows slower
// than the distance when a face moves outward (c < 1), so the test against the PARENT's box covers the whole chain.  Flat
// parents (zero thickness: invisible in the reference, aabb.rs:73) can never pass.  A failure proves nothing: the ray is
// simply traced again on the reference tree.
__device__ __forceinline__ bool ref_visible(const SceneDev& S, uint32_t ref, float3 o, float3 d, float t) {
    const uint32_t item = ref & RT_REF_ITEM_MASK;
    const float4 lo = __ldg(S.vis_box + 2u * item), hi = __ldg(S.vis_box + 2u * item + 1u);
    const float c = 4e-7f;
    const float px = o.x + t * d.x, py = o.y + t * d.y, pz = o.z + t * d.z;
    const float sx = fabsf(o.x) + fabsf(px), sy = fabsf(o.y) + fabsf(py), sz = fabsf(o.z) + fabsf(pz);
    return (px - lo.x > c * (fabsf(lo.x) + sx)) && (hi.x - px > c * (fabsf(hi.x) + sx)) &&
           (py - lo.y > c * (fabsf(lo.y) + sy)) && (hi.y - py > c * (fabsf(hi.y) + sy)) &&
           (pz - lo.z > c * (fabsf(lo.z) + sz)) && (hi.z - pz > c * (fabsf(hi.z) + sz));
}

// The exact form of the same proof, for the few hits (~5e-4 of the rays) the margins above cannot decide: walk the element's
// ancestors in the reference tree and evaluate, with the reference's own arithmetic (slab_axis: true IEEE divisions), whether
// t0_i <= t < t1_i on every axis of every ancestor box.  Given that t is the unique minimum over all elements, the reference's
// range end e is > t whenever it reaches an ancestor (e is inf or another element's accepted t), so min(t1_i, e) > t >=
// max(t0_i, start): Aabb::hit_cond (aabb.rs:65-78) passes for the whole chain, the element is reached and tested with a range
// that contains t, and it wins.  ~20 levels x 6 divisions instead of a second traversal on the reference tree (~300 rounds that
// the whole warp has to step through).  NaN (0/0) compares false: not proven.  A failure proves nothing: the ray is traced again.
__device__ __noinline__ bool ref_reaches_exact(const SceneDev& S, uint32_t ref, float3 o, float3 d, float t) {
    uint32_t node = __float_as_uint(__ldg(S.vis_box + 2u * (ref & RT_REF_ITEM_MASK)).w);
    for (uint32_t level = 0; level < 8192u; ++level) {
        if (node == 0xFFFFFFFFu) return true;  // walked past the root
        const float4 lo = __ldg(S.own_box + 2u * node), hi = __ldg(S.own_box + 2u * node + 1u);
        float t0, t1;
        slab_axis(lo.x, hi.x, o.x, d.x, t0, t1);
        if (!(t0 <= t && t < t1)) return false;
        slab_axis(lo.y, hi.y, o.y, d.y, t0, t1);
        if (!(t0 <= t && t < t1)) return false;
        slab_axis(lo.z, hi.z, o.z, d.z, t0, t1);
        if (!(t0 <= t && t < t1)) return false;
        node = __float_as_uint(lo.w);
    }
    return false;
}
__device__ __forceinline__ bool ref_proven(const SceneDev& S, uint32_t ref, float3 o, float3 d, float t) {
    return ref_visible(S, ref, o, d, t) || ref_reaches_exact(S, ref, o, d, t);
}

template <int MODE, bool COUNT, class KEYFN>
__device__ __forceinline__ bool trav_step(const SceneDev& S, Trav& T, uint2* stk, uint32_t stk_stride,
                                          TravCount* cnt, const KEYFN& keyfn) {
    return trav_round<MODE, COUNT>(S, T, (T.cur & RT_REF_LEAF) != 0u, stk, stk_stride, cnt, keyfn);
}

template <int MODE, bool COUNT = false, class KEYFN>
__device__ __forceinline__ HitRec traverse(const SceneDev& S, float3 o, float3 d, float time, float start, float end_in,
                                           uint2* stk, uint32_t stk_stride, const KEYFN& keyfn,
                                           TravCount* cnt = nullptr) {
    Trav T;
    HitRec best;
    best.ref = RT_NO_HIT;
    best.t = end_in;
    if (!trav_begin<MODE>(S, T, o, d, time, start, end_in)) return best;
    while (trav_step<MODE, COUNT>(S, T, stk, stk_stride, cnt, keyfn)) {}
    best.ref = T.best;
    best.t = T.end;
    return best;
}

// Whole-ray form of the fast path (rt_trace).  *proven = false -> the caller must trace the ray on the reference tree.
template <class KEYFN>
__device__ __forceinline__ HitRec traverse_fast(const SceneDev& S, float3 o, float3 d, float time, float start, float end_in, uint2* stk,
                                                uint32_t stk_stride, bool* proven, const KEYFN& keyfn) {
    Trav T;
    HitRec best;
    best.ref = RT_NO_HIT;
    best.t = end_in;
    *proven = false;
    trav_setup(S, T, o, d, time, start, end_in);
    if (!fast_eligible(T)) return best;
    if (!root_gate(S, o, d, start, end_in, false)) { *proven = true; return best; }  // the reference root test itself fails
    trav_setup_fast(S, T);
    while (trav_round_fast<false>(S, T, (T.cur & RT_REF_LEAF) != 0u, stk, stk_stride, nullptr, keyfn)) {}
    best.ref = T.best;
    best.t = T.end;
    *proven = T.best == RT_NO_HIT || (!(T.dpos & RT_TIE_BIT) && ref_proven(S, T.best, o, d, T.end));
    return best;
}

// ------------------------------------------------------------------------------------------------
// HitInteraction for the final hit of a segment (src/lib/hittable.rs:24-54 + the three geometry files)
// ------------------------------------------------------------------------------------------------
struct Interaction {
    float3 position, normal;
    float u, v;
    bool front_face;
    uint32_t material;
};
// Dir3::to_radian -> sphere uv  (vec3.rs:241-249, sphere_geometry.rs:56-59)
__device__ __forceinline__ void sphere_uv(float3 n, float& u, float& v) {
    float theta = acosf(n.y);
    float phi = atan2f(-n.z, n.x) + 3.14159265358979323846f;
    u = phi / 6.28318530717958647692f;
    v = theta / 3.14159265358979323846f;
}
__device__ __forceinline__ Interaction make_interaction(const SceneDev& S, uint32_t ref, float t, float3 o, float3 d, float time,
                                                        bool want_uv) {
    Interaction hi;
    uint32_t kind = (ref >> RT_REF_KIND_SHIFT) & 3u;
    uint32_t item = ref & RT_REF_ITEM_MASK;
    const float4* q = reinterpret_cast<const float4*>(S.prims + item);
    const float4* sq = reinterpret_cast<const float4*>(S.shade + item);
    WrapDev wr;
    const bool wrapped = (ref & RT_REF_WRAPPED) != 0u;
    if (wrapped) {  // the element sees the reverse-transformed ray (hittable.rs:234-244)
        wr = load_wrap(S, item);
        wrap_reverse_ray(wr, time, o, d);
    }
    float3 pos = add3(o, mul3(d, t));  // Ray::at, ray.rs:20-22
    float3 sn;
    hi.u = 0.0f; hi.v = 0.0f;
    float4 s3 = __ldg(sq + 3);
    hi.material = __float_as_uint(s3.w);
    if (wrapped && (wr.flags & 4u)) {  // VolumeGeometry::hit :334-341: normal UP, uv ZERO, front_face false ("arbitrary")
        hi.front_face = false;
        hi.normal = f3(0.0f, 1.0f, 0.0f);
        hi.position = pos;
        wrap_apply_hit(wr, time, hi.position, hi.normal);
        return hi;
    }
    if (kind == 0u) {
        float4 s = __ldg(q);
        sn = div3(sub3(pos, f3(s.x, s.y, s.z)), s.w);  // sphere_geometry.rs:44
        if (want_uv) sphere_uv(sn, hi.u, hi.v);
    } else if (kind == 2u) {
        PrimDev p;
        ldg_record(q, p.q0, p.q1, p.q2, p.q3);
        float3 p0 = f3(p.q0.x, p.q0.y, p.q0.z);
        float3 qq = sub3(pos, p0);
        float w1 = dot3(qq, f3(p.q1.z, p.q1.w, p.q2.x)) / p.q2.y;
        float w2 = dot3(qq, f3(p.q2.z, p.q2.w, p.q3.x)) / p.q3.y;
        float w0 = 1.0f - w1 - w2;
        float4 s0 = __ldg(sq), s1 = __ldg(sq + 1), s2 = __ldg(sq + 2);
        // math::interpolate: v0*w0 + v1*w1 + v2*w2, NOT re-normalised (Q8)   triangle_geometry.rs:31-32
        sn = f3(s0.x * w0 + s0.w * w1 + s1.z * w2, s0.y * w0 + s1.x * w1 + s1.w * w2, s0.z * w0 + s1.y * w1 + s2.x * w2);
        hi.u = s2.y * w0 + s2.w * w1 + s3.y * w2;
        hi.v = s2.z * w0 + s3.x * w1 + s3.z * w2;
    } else if (kind == 3u) {  // Aabb::hit, aabb.rs:92-101: face normal = signum(position - centre) on the hit plane's axis
        int plane = 0;
        const float4 mn = __ldg(q), mx = __ldg(q + 1);
        {   // re-run intersections_line (same arithmetic as the accepted test) to recover which face produced t
            const float mnv[3] = {mn.x - o.x, mn.y - o.y, mn.z - o.z}, mxv[3] = {mx.x - o.x, mx.y - o.y, mx.z - o.z}, dd[3] = {d.x, d.y, d.z};
            float near = -__int_as_float(0x7f800000), far = __int_as_float(0x7f800000);
            int np = 0, fp = 0;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const float t1 = mnv[i] / dd[i], t2 = mxv[i] / dd[i];
                const float t_min = fminf(t1, t2), t_max = fmaxf(t1, t2);
                if (t_min > near) { near = t_min; np = i; }
                if (t_max < far) { far = t_max; fp = i; }
            }
            plane = (near == t) ? np : fp;
        }
        const float centre = (comp3(f3(mx.x, mx.y, mx.z), plane) + comp3(f3(mn.x, mn.y, mn.z), plane)) * 0.5f;
        const float v = comp3(pos, plane) - centre;
        const float sg = (v != v) ? v : (signbit(v) ? -1.0f : 1.0f);  // f32::signum
        sn = f3(plane == 0 ? sg : 0.0f, plane == 1 ? sg : 0.0f, plane == 2 ? sg : 0.0f);
    } else {
        float4 a = __ldg(q), b = __ldg(q + 1);
        int p0, p1, n;
        rect_axes(__float_as_uint(b.y), p0, p1, n);
        // Q5: uv.y divides by (r0.hi - r1.lo)   rect_geometry.rs:43-46
        hi.u = (comp3(pos, p0) - a.y) / (a.z - a.y);
        hi.v = (comp3(pos, p1) - a.w) / (a.z - a.w);
        sn = f3(n == 0 ? -1.0f : 0.0f, n == 1 ? -1.0f : 0.0f, n == 2 ? -1.0f : 0.0f);
    }
    hi.front_face = dot3(sn, d) < 0.0f;  // hittable.rs:41
    hi.normal = hi.front_face ? sn : neg3(sn);
    hi.position = pos;
    if (wrapped) wrap_apply_hit(wr, time, hi.position, hi.normal);
    return hi;
}

// ------------------------------------------------------------------------------------------------
// Counter-based RNG + inversion samplers (must stay bit-identical to oracle/oracle_rng.hpp REPLAY mode)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2,
                                               uint32_t c3) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    return make_uint4(c0, c1, c2, c3);
}
__device__ __forceinline__ float u01(uint32_t v) { return (float)(v >> 8) * 5.9604644775390625e-8f; }

__device__ __forceinline__ void sincos_2pi(float u, float& s_out, float& c_out) {
    float x = u * 4.0f;
    int q = (int)x;
    float f = x - (float)q;
    float a = f * 1.57079632679489661923f;
    float a2 = a * a;
    float s = a * (1.0f + a2 * (-1.6666667163e-1f + a2 * (8.3333337680e-3f + a2 * (-1.9841270114e-4f +
              a2 * (2.7557314297e-6f + a2 * (-2.5052107943e-8f))))));
    float c = 1.0f + a2 * (-0.5f + a2 * (4.1666667908e-2f + a2 * (-1.3888889225e-3f + a2 * (2.4801587642e-5f +
              a2 * (-2.7557319224e-7f + a2 * (2.0876755880e-9f))))));
    switch (q & 3) {
        case 0: c_out = c;  s_out = s;  break;
        case 1: c_out = -s; s_out = c;  break;
        case 2: c_out = -c; s_out = -s; break;
        default: c_out = s; s_out = -c; break;
    }
}
__device__ __forceinline__ float cbrt_unit(float u) {
    if (!(u > 0.0f)) return 0.0f;
    float y = __uint_as_float(__float_as_uint(u) / 3u + 709921077u);
#pragma unroll
    for (int i = 0; i < 4; ++i) y = (2.0f * y + u / (y * y)) * 0.333333343267f;
    return y;
}
// deterministic natural log on [0,1), IEEE ops only — bit-identical to oracle_rng.hpp ln_unit
__device__ __forceinline__ float ln_unit(float u) {
    if (!(u > 0.0f)) return -__int_as_float(0x7f800000);
    uint32_t b = __float_as_uint(u);
    int e = (int)(b >> 23) - 127;
    float m = __uint_as_float((b & 0x007FFFFFu) | 0x3F800000u);
    if (m > 1.41421356f) { m = m * 0.5f; e += 1; }
    float s = (m - 1.0f) / (m + 1.0f);
    float s2 = s * s;
    float p = 2.0f * s * (1.0f + s2 * (0.333333343f + s2 * (0.2f + s2 * (0.142857149f + s2 * 0.111111112f))));
    return (float)e * 0.693147182f + p;
}
__device__ __forceinline__ float3 sample_unit_sphere(float ua, float ub) {  // stands in for rand_distr::UnitSphere
    float z = 1.0f - 2.0f * ua;
    float r2 = 1.0f - z * z;
    float r = sqrtf(r2 > 0.0f ? r2 : 0.0f);
    float s, c;
    sincos_2pi(ub, s, c);
    return f3(r * c, r * s, z);
}

// Camera::ray (src/lib/camera.rs:175-200) for the sample whose Philox block 0 is `r`
__device__ __forceinline__ void camera_ray(const CameraDev& C, float px, float py, uint4 r, float3& o, float3& d) {
    float3 offset = f3(0.0f, 0.0f, 0.0f);
    if (C.lens_radius > 0.0f) {
        float rad = sqrtf(u01(r.z));  // stands in for rand_distr::UnitDisc
        float s, c;
        sincos_2pi(u01(r.w), s, c);
        float rdx = rad * c, rdy = rad * s;
        offset = mul3(add3(mul3(C.unit_right, rdx), mul3(C.unit_up, rdy)), C.lens_radius);
    }
    o = add3(C.position, offset);
    d = unit3(sub3(sub3(add3(C.ulc, mul3(C.scaled_right, px)), mul3(C.scaled_up, py)), offset));
}

// f32::sin for the procedural textures (texture.rs:32, 50).  CUDA's sinf carries a Payne-Hanek slow path (|x| > 105 615) that works
// on a local-memory array: never executed for scene-scale arguments, but ~500 LDL / STL instructions in every kernel that samples a
// texture.  This version reduces the argument in double precision (two-term Cody-Waite with fused multiply-adds: exact to ~1e-16
// for |x| < 1e15), evaluates the quadrant's polynomial in double and rounds once: within 1 ulp of the correctly rounded value that
// glibc (the oracle, and Rust's std on Linux) returns — the same tolerance the tests already grant device sinf.
__device__ __forceinline__ float sin_f32(float x) {
    const double xd = (double)x;
    const double k = rint(xd * 0.63661977236758134308);  // nearest multiple of pi/2
    double r = fma(-k, 1.57079632679489655800e+00, xd);
    r = fma(-k, 6.12323399573676603587e-17, r);
    const long long q = __double2ll_rn(k);
    const double r2 = r * r;
    const double sn = r * fma(r2, fma(r2, fma(r2, fma(r2, fma(r2, fma(r2, 1.6059043836821613e-10, -2.5052108385441720e-08), 2.7557319223985893e-06),
                                                                 -1.9841269841269841e-04), 8.3333333333333333e-03), -1.6666666666666666e-01), 1.0);
    const double cs = fma(r2, fma(r2, fma(r2, fma(r2, fma(r2, fma(r2, fma(r2, -1.1470745597729725e-11, 2.0876756987868100e-09), -2.7557319223985888e-07),
                                                                 2.4801587301587302e-05), -1.3888888888888889e-03), 4.1666666666666664e-02), -0.5), 1.0);
    const double v = (q & 1) ? cs : sn;
    return (float)((q & 2) ? -v : v);  // NaN / inf arguments: r is NaN, so is the result (as f32::sin)
}

// Texture::sample (src/lib/texture.rs:23-43)
__device__ __forceinline__ uint32_t sat_u32(float v) {  // Rust `as u32` (saturating, NaN -> 0)
    if (!(v == v)) return 0u;
    if (v <= 0.0f) return 0u;
    if (v >= 4294967296.0f) return 0xFFFFFFFFu;
    return (uint32_t)v;
}
// Perlin::noise / ::turbulence (src/lib/perlin.rs:37-91), bits = 8
__device__ __forceinline__ float perlin_noise(const TextureDev& t, float3 p) {
    const float fx = floorf(p.x), fy = floorf(p.y), fz = floorf(p.z);
    const float u = p.x - fx, v = p.y - fy, w = p.z - fz;
    const int i = (int)fx, j = (int)fy, k = (int)fz;
    const float uu = u * u * (3.0f - 2.0f * u), vv = v * v * (3.0f - 2.0f * v), ww = w * w * (3.0f - 2.0f * w);
    float accum = 0.0f;
#pragma unroll
    for (int di = 0; di < 2; ++di)
#pragma unroll
        for (int dj = 0; dj < 2; ++dj)
#pragma unroll
            for (int dk = 0; dk < 2; ++dk) {
                const uint32_t r = t.perm[(i + di) & 255] ^ t.perm[256 + ((j + dj) & 255)] ^ t.perm[512 + ((k + dk) & 255)];
                const float3 c = f3(t.ranvec[3 * r], t.ranvec[3 * r + 1], t.ranvec[3 * r + 2]);
                const float fi = (float)di, fj = (float)dj, fk = (float)dk;
                const float3 weight = f3(u - fi, v - fj, w - fk);
                accum += (fi * uu + (1.0f - fi) * (1.0f - uu)) * (fj * vv + (1.0f - fj) * (1.0f - vv)) * (fk * ww + (1.0f - fk) * (1.0f - ww)) *
                         dot3(c, weight);
            }
    return accum;
}
__device__ __forceinline__ float perlin_turbulence(const TextureDev& t, float3 p, int depth, float fall_off) {
    float accum = 0.0f, weight = 1.0f;
    for (int d = 0; d < depth; ++d) {
        accum += weight * perlin_noise(t, p);
        weight *= fall_off;
        p = mul3(p, 2.0f);
    }
    return fabsf(accum);
}
// does sampling this texture (through any checker nesting) read uv?
__device__ __forceinline__ bool texture_needs_uv(const SceneDev& S, uint32_t tex) {
    const TextureDev t = S.textures[tex];
    if (t.kind == 1u) return true;
    if (t.kind == 2u) return true;  // a checker may select an image texture: be conservative
    return false;
}
__device__ __forceinline__ float3 texture_sample(const SceneDev& S, uint32_t tex, float u, float v, float3 position) {
    TextureDev t = S.textures[tex];
    for (int depth = 0; depth < 8 && t.kind == 2u; ++depth) {  // Checker (texture.rs:26-35): pick even / odd by the sign of the sine product
        const float sx = position.x * t.inv_frequency, sy = position.y * t.inv_frequency, sz = position.z * t.inv_frequency;
        const float sines = sin_f32(sx) * sin_f32(sy) * sin_f32(sz);
        t = S.textures[sines < 0.0f ? t.even : t.odd];
    }
    if (t.kind == 0u) return f3(t.r, t.g, t.b);
    if (t.kind == 3u) {  // Marble (texture.rs:44-51): 0.5 * (1 + sin(scale * z + 10 * turbulence(p, 7, 0.5)))
        const float v = 0.5f * (1.0f + sin_f32(position.z * t.inv_frequency + 10.0f * perlin_turbulence(t, position, 7, 0.5f)));
        return f3(v, v, v);
    }
    uint32_t pu = min(sat_u32(u * (float)t.width), t.width - 1u);
    uint32_t pv = min(sat_u32(v * (float)t.height), t.height - 1u);
    const uint8_t* px = t.texels + 3ull * ((size_t)pv * t.width + pu);
    return f3((float)px[0] / 255.0f, (float)px[1] / 255.0f, (float)px[2] / 255.0f);  // color.rs:29-35
}

}  // namespace rt
