/* b200rt.h — C ABI of the B200-native render hot path (libb200rt.so).
 *
 * Drop-in boundary for the reference's `src/lib` render path (martingleich/RayTracingInAWeekend):
 * a host layer flattens the reference scene graph into the POD arrays below and calls these entry
 * points; everything behind them is hand-written sm_100a CUDA.  Plain pointers and sizes only.
 *
 * Reference interfaces replaced (file:line relative to the reference repo):
 *   rt_scene_create   <- the data reachable from `World` (src/lib/rendering.rs:12-17): `Scene::new(root)`
 *                        (src/lib/hittable.rs:125-128) with root = the flat `BoundingVolumeHierarchy`
 *                        built by `BoundingVolumeHierarchy::new` (hittable.rs:360-380) from the element list of
 *                        `NodeRef::finish` (src/app/worlds/world_builder.rs:273-328); `Camera` (src/lib/camera.rs:154-164),
 *                        `BackgroundColor` (src/lib/background_color.rs:3-6), `WorldScatteringDistributionProvider::Rect`
 *                        (src/lib/world_scattering_distribution.rs:4-9), `Material` (src/lib/material.rs:43-49),
 *                        `Texture` (src/lib/texture.rs:4-20).
 *   rt_trace          <- `Scene::hit(&self, &Ray, &Range<f32>, &mut TRng) -> Option<HitInteraction>` (hittable.rs:130-137)
 *   rt_render         <- `render(image_size, thread_count, samples_per_pixel, max_depth, &World, RenderMode) -> Vec<Color>`
 *                        (src/lib/rendering.rs:121-128)
 *   rt_render_planes_device / rt_film_scale_device
 *                     <- the per-thread planes + `merge_planes` of the same function (rendering.rs:161-210, 239-252),
 *                        split so that one process per GPU renders a subset of the sample lanes and the film is
 *                        combined with an NCCL all-reduce in between.
 *   rt_to_rgb8_gamma2 <- `Color::to_rgb8_gamma2` (src/lib/color.rs:43-48), applied by the caller in src/app/main.rs:53-56.
 *
 * Every function returns 0 on success or a negative RT_E* code; rt_last_error() returns a thread-local
 * message.  Nothing here ever falls back to a CPU implementation: without a CUDA device the calls fail.
 */
#ifndef B200RT_H
#define B200RT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RT_OK 0
#define RT_EINVAL (-1)  /* malformed scene / arguments                        */
#define RT_ECUDA (-2)   /* CUDA runtime error (no device, launch failure ...) */
#define RT_ENOMEM (-3)  /* host or device allocation failed                   */
#define RT_EUNSUPPORTED (-4) /* not supported here: rt_render_multi without peer access between the GPUs.  (Scenes: every NodeRef::finish output has a flat form; Group / nested BVH / unbounded items are rejected by the flattener on the caller side) */

#define RT_LEAF_BIT 0x80000000u /* BVH child reference: RT_LEAF_BIT | item index  (reference: usize::MAX - index) */
#define RT_MISS 0xFFFFFFFFu     /* hit id of a ray that hit nothing                */

/* Geometry kinds (reference enum `Geometry`, hittable.rs:104-110) */
enum { RT_GEO_SPHERE = 0, RT_GEO_RECT = 1, RT_GEO_TRIANGLE = 2, RT_GEO_BOX = 3 };
/* Run-time wrappers of a scene element (SceneElement::Transformation / ::Animation, hittable.rs:117-118), applied in the
 * order NodeRef::finish nests them: Animation(Transformation(SurfaceGeometry))  (world_builder.rs:310-315) */
enum { RT_WRAP_TRANSFORMATION = 1, RT_WRAP_ANIMATION = 2,
       RT_WRAP_VOLUME = 4 /* SceneElement::VolumeGeometry (hittable.rs:116, 294-342): geom is the boundary, material the phase function */ };
/* RectPlane (rect_geometry.rs:8-22) */
enum { RT_RECT_XY = 0, RT_RECT_XZ = 1, RT_RECT_YZ = 2 };
/* Material kinds (material.rs:43-49) */
enum { RT_MAT_LAMBERT = 0, RT_MAT_METAL = 1, RT_MAT_DIELECTRIC = 2, RT_MAT_DIFFUSE_LIGHT = 3, RT_MAT_ISOTROPIC = 4 };
/* Texture kinds (texture.rs:4-20) */
enum { RT_TEX_SOLID = 0, RT_TEX_IMAGE = 1, RT_TEX_CHECKER = 2, RT_TEX_MARBLE = 3 };
/* RenderMode (rendering.rs:94-98) */
enum { RT_MODE_DEFAULT = 0, RT_MODE_NORMALS = 1 };

/* One flattened scene element = SceneElement::SurfaceGeometry(geo, material).  Its index in the array is the
 * "hit id" (position in `all_elements`, world_builder.rs:316).  `geom` holds, by kind:
 *   sphere   : center xyz, radius                               (sphere_geometry.rs:6-9)
 *   rect     : (float)plane, dist, r0.lo, r0.hi, r1.lo, r1.hi   (rect_geometry.rs:24-30)
 *   triangle : positions[3] xyz, normals[3] xyz, texture_coords[3] uv   (triangle_geometry.rs:6-10)
 *   box      : min xyz, max xyz                                 (Geometry::AxisAlignedBox(Aabb), aabb.rs:10-13)
 * wrap_flags / transform / velocity describe the optional run-time wrappers: Transformation{offset, y_sine, y_cosine}
 * (transformations.rs:4-8) and Animation(velocity) (hittable.rs:117). */
typedef struct RtItem {
    uint32_t kind;
    uint32_t material;
    float geom[24];
    uint32_t wrap_flags;  /* RT_WRAP_* bits */
    float transform[5];   /* offset xyz, y_sine, y_cosine */
    float velocity[3];
    float neg_inv_density; /* RT_WRAP_VOLUME: -1 / density (VolumeGeometry::new, hittable.rs:301-307) */
} RtItem;

/* BoundingVolumeNode (hittable.rs:345-351) in the reference's pre-order numbering. */
typedef struct RtBvhNode {
    float aabb_min[3];
    float aabb_max[3];
    uint32_t left;  /* node index, or RT_LEAF_BIT | item */
    uint32_t right;
    uint32_t axis_id;
} RtBvhNode;

typedef struct RtMaterial {
    uint32_t kind;
    uint32_t texture; /* albedo / emit; ignored for dielectric */
    float param;      /* metal: fuzz; dielectric: index_of_refraction */
} RtMaterial;

typedef struct RtTexture {
    uint32_t kind;
    uint32_t image;      /* RT_TEX_IMAGE: index into images; RT_TEX_MARBLE: index into perlins */
    float rgb[3];        /* RT_TEX_SOLID */
    float inv_frequency; /* RT_TEX_CHECKER (texture.rs:8-12); RT_TEX_MARBLE: scale (texture.rs:13-16) */
    uint32_t even, odd;  /* RT_TEX_CHECKER: texture indices (may themselves be checkers, nesting <= 8) */
} RtTexture;

typedef struct RtImage {
    uint32_t width, height;
    const uint8_t* rgb8; /* row 0 = top, 3 bytes per texel (image::RgbImage) */
} RtImage;

/* Perlin noise tables (perlin.rs:10-16) with bits = 8: 256 gradient vectors + three 256-entry permutations. */
typedef struct RtPerlin {
    const float* ranvec;   /* 256 * 3 */
    const uint32_t* perm_x;
    const uint32_t* perm_y;
    const uint32_t* perm_z;
} RtPerlin;

typedef struct RtCamera { /* camera.rs:154-164, private fields included */
    float position[3];
    float upper_left_corner[3];
    float unit_right[3];
    float unit_up[3];
    float scaled_right[3];
    float scaled_up[3];
    float lens_radius;
    float time_start, time_end; /* motion_blur(start, end): ray.time = U[start, end) per camera sample (camera.rs:185-190) */
    float shutter_pace[2];
} RtCamera;

typedef struct RtSceneDesc {
    uint32_t n_items;
    const RtItem* items;
    uint32_t n_nodes;       /* n_items - 1 (0 for a single item) */
    const RtBvhNode* nodes; /* reference topology; node 0 first */
    uint32_t root;          /* initial_index: 0, or RT_LEAF_BIT|0 for a single item */
    uint32_t n_materials;
    const RtMaterial* materials;
    uint32_t n_textures;
    const RtTexture* textures;
    uint32_t n_images;
    const RtImage* images;
    uint32_t n_perlins;
    const RtPerlin* perlins;
    RtCamera camera;
    uint32_t background_is_sky; /* BackgroundColor::Sky vs ::Solid */
    float background_rgb[3];
    uint32_t has_light;         /* Option<WorldScatteringDistributionProvider::Rect> */
    uint32_t light_plane;
    float light_dist, light_r0[2], light_r1[2];
} RtSceneDesc;

typedef struct RtRenderParams {
    int32_t width, height;      /* Size2i */
    uint64_t samples_per_pixel;
    int32_t max_depth;
    uint32_t thread_count;      /* reference: worker threads; here: sample lanes in flight per pixel.
                                   Lane id gets spp/T (+1 if id < spp%T) samples — split_work_tasks, rendering.rs:222-237 */
    uint32_t render_mode;       /* RT_MODE_* */
    uint64_t seed;              /* Philox key (reference seeds from OS entropy, rendering.rs:160) */
} RtRenderParams;

typedef struct RtStats {
    uint64_t ray_segments;   /* top-level closest-hit queries (rendering.rs:25) */
    uint64_t paths;          /* camera samples generated */
    uint64_t iterations;     /* wavefront iterations (extend+shade pairs) */
    uint64_t kernel_launches;
    double render_ms;        /* device time of the render loop (CUDA events), excluding upload/download */
    double extend_ms;        /* device time spent in the extend (closest-hit) kernel, when profiling is on */
    double shade_ms;
    uint64_t node_visits;    /* BVH node fetches (64 B each) — only with rt_set_option("count", 1) */
    uint64_t prim_tests;     /* primitive intersection tests    — only with rt_set_option("count", 1) */
    uint64_t exact_steps;    /* node steps that fell back to true divisions — only with rt_set_option("count", 1) */
    uint64_t redo_rays;      /* traversal 3: rays the fast topology could not prove, traced again on the reference tree — only with "count" */
    uint32_t max_stack;      /* deepest traversal stack use of any ray (entries) — only with "count" */
    uint32_t stack_capacity; /* entries per thread the launch provided (tree depth + 2): max_stack must never exceed it */
    uint32_t slots;          /* path slots in flight the library chose: min("max_slots", lanes of this rank x pixels) */
    uint32_t pipes;          /* wavefront pipelines (streams) used */
    double merge_ms;         /* rt_render_multi: device time of this GPU's share of the film combine (k_merge_peers) */
    uint64_t warp_rounds, busy_lane_rounds; /* only with "count": scheduling rounds of the extend warps and lanes holding a ray in them */
    uint64_t rounds_hist[24];/* only with "count": rays by floor(log2(traversal rounds)) — the tail that sets a launch's duration */
} RtStats;

typedef struct RtScene RtScene;

const char* rt_last_error(void);
int rt_device_count(int* out_count);

/* Copies everything; the caller keeps ownership of its buffers.  `device` = CUDA ordinal. */
int rt_scene_create(const RtSceneDesc* desc, int device, RtScene** out_scene);
void rt_scene_destroy(RtScene* scene);
/* Device bytes held by the packed scene (nodes, primitives, shading records, tables, images). */
int rt_scene_device_bytes(const RtScene* scene, uint64_t* out_bytes);
/* Test aid for row f4: the accelerator tree of the scene as the kernels read it.  *out_n = its node count (0: the scene has
 * none: fewer than "rebuild_min_items" elements, or rebuilding is off), *out_root = the root reference, *out_depth = its depth.
 * out_nodes64 (may be NULL): room for cap_nodes 64-byte records = 16 words each: the left child's box (min xyz, max xyz), the
 * right child's box, then left reference, right reference, split axis, packed guard band.  A reference is a node index, or has
 * bit 31 set and the element index in its low 28 bits. */
int rt_scene_accelerator_nodes(const RtScene* scene, void* out_nodes64, uint64_t cap_nodes, uint64_t* out_n, uint32_t* out_root, uint32_t* out_depth);

/* Closest hit for n caller rays (host buffers).  Volumes draw one random number per tested volume element inside the traversal
 * (the reference passes an rng into Scene::hit for this): rt_trace keys that draw by (ray index, element index), seed 0.  org/dir: n*3 floats; outputs: hit_id n, t n, normal n*3,
 * front_face n, uv n*2 (any output pointer may be NULL).  Half-open range [t_min, t_max).  Directions should be unit vectors as
 * Ray::new demands (ray.rs:13); a ray whose |dir|^2 is off by more than 1e-5 is still answered with the reference's arithmetic
 * (which is not the geometric intersection for spheres then), using the reference node rule alone. */
int rt_trace(RtScene* scene, const float* org_xyz, const float* dir_xyz, uint64_t n, float t_min, float t_max,
             uint32_t* hit_id, float* t, float* normal_xyz, uint8_t* front_face, float* uv);
/* Ray::time of the rays handed to rt_trace / rt_trace_device (only Animation leaves look at it; default 0). */
int rt_scene_set_ray_time(RtScene* scene, float time);
/* Same on device buffers (as allocated by the caller, e.g. a torch tensor's data_ptr) on `stream`. */
int rt_trace_device(RtScene* scene, const float* d_org_xyz, const float* d_dir_xyz, uint64_t n, float t_min, float t_max,
                    uint32_t* d_hit_id, float* d_t, float* d_normal_xyz, uint8_t* d_front_face, float* d_uv,
                    void* cuda_stream);

/* Whole render into a host buffer: out_rgb = width*height*3 floats, row 0 = top, linear RGB. */
int rt_render(RtScene* scene, const RtRenderParams* params, float* out_rgb, RtStats* out_stats);
/* Whole render into a device buffer (width*height*3 floats). */
int rt_render_device(RtScene* scene, const RtRenderParams* params, float* d_out_rgb, void* cuda_stream,
                     RtStats* out_stats);

/* rt_render_device plus the per-pixel second moment: d_out_sumsq[p] = sum over all samples of (r^2 + g^2 + b^2) (width*height
 * floats).  With the film (the mean) this gives the per-sample variance the image-RMSE acceptance bound is derived from
 * (RMSE <= 1.25 sqrt(2 mean(sigma^2) / N)); the reference has no counterpart (rendering.rs:172-179 keeps only the sum). */
int rt_render_moments_device(RtScene* scene, const RtRenderParams* params, float* d_out_rgb, float* d_out_sumsq, void* cuda_stream,
                             RtStats* out_stats);

/* Multi-GPU split of render(): this rank renders sample lanes {lane_first, lane_first+lane_stride, ...} < T and
 * writes  sum over its lanes of (lane plane / lane sample count)  into d_plane_sum (width*height*3 floats).
 * After summing d_plane_sum over ranks (NCCL all-reduce), rt_film_scale_device multiplies by 1/T (merge_planes). */
int rt_render_planes_device(RtScene* scene, const RtRenderParams* params, uint32_t lane_first, uint32_t lane_stride,
                            float* d_plane_sum, void* cuda_stream, RtStats* out_stats);
int rt_film_scale_device(float* d_rgb, uint64_t n_floats, float multiplier, void* cuda_stream);

/* render() on n_scenes GPUs of one box in ONE blocking call (replaces the worker-thread scope + merge_planes of rendering.rs:161-210,
 * 239-252 across devices).  scenes[g] = the same RtSceneDesc created on n_scenes different devices.  GPU g renders sample lanes
 * g, g + n_scenes, ... of the thread_count lanes (one host thread per GPU); the film is combined inside the library by a
 * fixed-order reduce over NVLink peer memory fused with the finalize step (lane T-1 first, then ascending, x 1/T — merge_planes'
 * own order), so out_rgb has the SAME BITS for every n_scenes, n_scenes = 1 and rt_render included.  Needs peer access between
 * the GPUs (RT_EUNSUPPORTED otherwise).  out_stats: NULL or n_scenes entries (per GPU).
 * rt_render_multi_device leaves the complete film on every GPU instead (d_out_rgb[g] on scenes[g]'s device; the all-gather is
 * fused into the reduce as peer stores). */
int rt_render_multi(RtScene* const* scenes, int n_scenes, const RtRenderParams* params, float* out_rgb, RtStats* out_stats);
int rt_render_multi_device(RtScene* const* scenes, int n_scenes, const RtRenderParams* params, float* const* d_out_rgb, RtStats* out_stats);

/* Color::to_rgb8_gamma2 on device data. */
int rt_to_rgb8_gamma2_device(const float* d_rgb, uint64_t n_pixels, uint8_t* d_out_rgb8, void* cuda_stream);

/* Knobs (process-wide; every rt_scene_create / render call reads ONE consistent snapshot at entry, rt_set_option is
 * serialised with those reads, so the calls stay re-entrant).  "max_slots": path slots in flight per render (default 160 Mi).
 * The render's work is thread_count x pixels CHAINS (all samples of one lane at one pixel, rendered in order by one slot, so every
 * per-lane sum keeps the reference's order); a slot whose chain has ended takes the next chain, so the wavefronts stay full and the
 * path-state workspace (120 B per slot) does not grow with thread_count.  The film is bit-identical for any value.
 * "traversal": 0 = reference node test only, true divisions (the reference's algorithm, for A/B
 * parity sweeps); 1 = + tight carried-interval cull; 2 = + reciprocal fast path with exact fallback; 3 (default) = a
 * rebuilt (SAH) topology over the same elements finds the closest accepted element, the result is kept only when the
 * reference traversal provably returns the same element: the hit point lies strictly inside the element's parent box of the
 * reference tree (margin test), or — evaluated with the reference's own divisions — inside every ancestor box (exact walk); two
 * elements that accept exactly the same t are ordered by the reference's visit rank for the ray's direction octant (plain scenes).
 * The proof runs in the shade kernel; what it cannot decide goes through the queue again and is traced with 2 on the reference
 * tree (RtStats.redo_rays; ~1e-5 of the rays); scenes with
 * fewer than "rebuild_min_items" (32) elements use 2, as does rt_trace on scenes with Animation leaves (the rebuilt tree bounds
 * moving elements over the camera's shutter interval only).  "rebuild": 0 = do not build the second topology at rt_scene_create;
 * "rebuild_wrapped": 0 = not for scenes with wrapped elements / volumes.
 * "rebuild_depth_extra": depth cap of the rebuilt tree above the reference depth (default 4).  "presplit": reference budget of
 * the rebuilt tree in per cent of the element count (default 800; <= 100 = one reference per element): sliver triangles are
 * cut into several references with tighter boxes (an element may sit in several leaves of the accelerator tree).
 * "presplit_device": 1 (default) = that cutting runs on the device for scenes of >= 16 Ki elements, 0 = on the host cores (same
 * set of references either way).  "build_device": 1 (default) = after a device pre-split the accelerator tree is built on the device too
 * (rt_sahbuild.cuh: nothing comes back to the host), 0 = by the host builder.  "presplit_large": the budget (per cent) for scenes
 * above 150 000 elements (default 200).  "device_cache_mb": device memory of destroyed scenes (and of the hand-off's scratch) kept per GPU for the
 * next rt_scene_create, in MB (default 2048; 0 = every release is a cudaFree; rt_shutdown gives everything back).
 * "count": 1 = maintain node_visits / prim_tests / exact_steps.  "profile": 1 = time extend / shade separately.
 * "refill": busy-lane threshold of the persistent extend warps.  "leaf_weight": phase-vote weight.  "check_every": host
 * polling period (iterations).  "pipes": independent wavefront pipelines (streams) per render (default 0 = auto, 3 for >= 64 Mi path slots, else 4: one
 * pipeline's launch tails overlap the others' work; the film is bit-identical for any value).  "tail_budget": parked traversals —
 * an extend warp that can no longer refill gives its rays max(this floor, 4 x the previous launch's mean tail) more rounds, then
 * parks what is still running (cursor + stack + best hit); the path waits one iteration and a second, small launch of the extend
 * kernel finishes the segment beside the next queue launch (default 64; 0 = never park; -n = exactly n rounds, for tests).
 * "blocks_per_sm": cap on resident blocks of the persistent kernels.  "carveout": preferred shared-memory carve-out of the extend kernel in
 * per cent (default 75: 171 KB of the SM's 228 KB for the traversal stacks, the rest stays L1; -1 = driver default).
 * "debug_skip_redo": 1 = measurement only, keeps fast-path hits that were not proven (films are then NOT the reference's).
 * Environment: RT_ITER_LOG=<path> dumps (pipe, iteration, live paths, extend ms, shade ms, counters, parked, budget) of the last
 * render as CSV; RT_LONG_RAYS=<path> (with "count") the rays that took >= 4096 traversal rounds; RT_TAIL_LOG=1 (with "count") the
 * share of warp rounds after the queue ran dry / with a lane on the reference tree; RT_PACK_LOG=1 the phase times of
 * rt_scene_create. */
int rt_set_option(const char* key, int64_t value);

/* Test aid for row f4's pre-split (rt_rebuild.hpp / rt_presplit.cuh), the step of rt_scene_create that cuts sliver triangles into
 * several references for the accelerator tree (the reference has no counterpart: its tree, hittable.rs:382-427, holds every
 * element once).  boxes6: n x (min xyz, max xyz); tri9: n x (p0, p1, p2), read where is_tri != 0; budget: references allowed.
 * device < 0: the host implementation; otherwise the device one on that GPU.  out_items / out_boxes6: room for max(budget, n)
 * references; *out_n: how many were made.  Both implementations produce the same SET of references (their order differs). */
int rt_presplit_refs(const float* boxes6, const float* tri9, const uint8_t* is_tri, uint32_t n, uint64_t budget, int device,
                     uint32_t* out_items, float* out_boxes6, uint64_t* out_n);

/* Frees the per-device render workspaces (path-state queues, per-lane accumulators, streams, events) that the library keeps
 * between renders.  Optional; call it only when no render is in flight.  The next render allocates them again. */
int rt_shutdown(void);

#ifdef __cplusplus
}
#endif
#endif /* B200RT_H */
