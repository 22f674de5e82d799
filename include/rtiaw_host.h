/* rtiaw_host.h — C shim over the C++ host mirror of the reference API (raytracinginaweekend_b200/host/rtiaw.hpp),
 * exported by librtiaw_host.so.  It exists so that non-C++ callers (the Python tests / bench, a Rust `extern "C"` block)
 * can drive the same call sequence as src/app/main.rs: build a world -> render() -> gamma.  Everything that touches rays
 * goes through libb200rt.so (include/b200rt.h); this library only authors / flattens scenes on the host.
 *
 * Reference interfaces mirrored (file:line in the reference repo):
 *   rtw_world_create          <- `create_world_*(&wb, &mut rng) -> World` selected by name (src/app/main.rs:25-37,
 *                                src/app/worlds/demo_worlds.rs) incl. `NodeRef::finish` (world_builder.rs:273-290)
 *   rtw_world_create_elements <- `Scene::new(BoundingVolumeHierarchy::new(elements))` + `World{..}` (hittable.rs:125-128,360-380;
 *                                rendering.rs:12-17) from caller-supplied SurfaceGeometry elements
 *   rtw_world_export          <- read-back of the flattened scene (element list, BVH nodes, tables, camera)
 *   rtw_render                <- `render(image_size, thread_count, samples_per_pixel, max_depth, &world, render_mode)` (rendering.rs:121-128)
 *   rtw_hit                   <- `Scene::hit` (hittable.rs:130-137), batched
 *   rtw_load_obj              <- `load_obj_mesh` (src/app/obj_loader.rs:7-22)
 * All functions return 0 / a handle on success, non-zero / NULL on failure with the message in rtw_last_error().
 */
#ifndef RTIAW_HOST_H
#define RTIAW_HOST_H
#include <stdint.h>

#include "b200rt.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct RtwWorld RtwWorld;

const char* rtw_last_error(void);

RtwWorld* rtw_world_create(const char* name, const char* suzanne_obj_path, const char* cube_obj_path, uint32_t earth_w,
                           uint32_t earth_h, const uint8_t* earth_rgb8);
/* geom: 24 floats per element (layout of RtItem.geom); cam: position, look-at basis as the 23 floats of RtCamera in order;
 * bg: is_sky, r, g, b; light: has, plane, dist, r0lo, r0hi, r1lo, r1hi;
 * wrap10 (optional): per element flags (RT_WRAP_*), offset xyz, y_sine, y_cosine, velocity xyz, -1/density. */
RtwWorld* rtw_world_create_elements(uint32_t n, const uint32_t* kinds, const uint32_t* materials_of, const float* geom,
                                    uint32_t n_materials, const RtMaterial* materials, uint32_t n_textures,
                                    const RtTexture* textures, uint32_t n_images, const RtImage* images, const float* cam23,
                                    const float* bg4, const float* light7, const float* wrap10 /* may be NULL */,
                                    uint32_t n_perlins, const RtPerlin* perlins);
void rtw_world_destroy(RtwWorld* w);

/* info[11]: n_items, n_nodes, root, depth, default_w, default_h, n_materials, n_textures, n_images, has_light, n_perlins */
int rtw_world_info(const RtwWorld* w, uint64_t* info);
/* Copies the flattened scene into caller arrays sized from rtw_world_info (same layouts as RtItem / RtBvhNode / ...). */
int rtw_world_export(const RtwWorld* w, RtItem* items, RtBvhNode* nodes, RtMaterial* materials, RtTexture* textures,
                     uint32_t* image_wh, float* cam23, float* bg4, float* light7);
const uint8_t* rtw_world_image(const RtwWorld* w, uint32_t index);
/* Perlin table `index`: 768 gradient floats, then perm_x | perm_y | perm_z (256 u32 each). */
int rtw_world_perlin(const RtwWorld* w, uint32_t index, float* ranvec768, uint32_t* perm768);

/* The device-resident scene (created on first use; owned by the world). */
int rtw_world_device(RtwWorld* w, int device, RtScene** out_scene);
/* Drop the device copy (the next use uploads the scene again: lets callers time the host->device hand-off). */
int rtw_world_release_device(RtwWorld* w);
/* Host bytes that rt_scene_create copies for this world (items, nodes, tables, images). */
int rtw_world_upload_bytes(const RtwWorld* w, uint64_t* out_bytes);

int rtw_render(RtwWorld* w, int32_t width, int32_t height, uint64_t thread_count, uint64_t samples_per_pixel,
               int32_t max_depth, uint32_t render_mode, uint64_t seed, int device, float* out_rgb, RtStats* out_stats);
int rtw_hit(RtwWorld* w, uint64_t n, const float* org_xyz, const float* dir_xyz, float t_min, float t_max, int device,
            uint32_t* hit_id, float* t, float* normal_xyz, uint8_t* front_face, float* uv);

/* out24: 24 floats per triangle (positions, normals, uvs); returns the triangle count in *n_tris (even if > cap). */
int rtw_load_obj(const char* path, uint64_t* n_tris, float* out24, uint64_t cap);

#ifdef __cplusplus
}
#endif
#endif
