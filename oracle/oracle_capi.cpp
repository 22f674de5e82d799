// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_core.hpp header).  PARITY UNPINNED.
//
// C entry points (ctypes) over the restatement: world construction, flat export of the flattened
// scene (for structural parity with the product's flattener), closest-hit queries (Scene::hit,
// src/lib/hittable.rs:130-137) and render() (src/lib/rendering.rs:121-220) with the reference's
// thread-per-sample-split decomposition.
#include <chrono>
#include <cstring>
#include <string>
#include <thread>

#include "oracle_app.hpp"

using namespace orc;

namespace {
thread_local std::string g_err;
struct Handle {
    std::unique_ptr<BuiltWorld> bw;
};
const uint32_t LEAF_BIT = 0x80000000u;
uint32_t enc_child(size_t c, size_t n_items) { return c > n_items ? (LEAF_BIT | (uint32_t)(USIZE_MAX - c)) : (uint32_t)c; }
}  // namespace

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }

void* orc_world_create(const char* name, const char* suzanne_obj, const char* cube_obj, uint32_t earth_w,
                       uint32_t earth_h, const uint8_t* earth_rgb) {
    try {
        SceneAssets a;
        a.suzanne_obj = suzanne_obj ? suzanne_obj : "";
        a.cube_obj = cube_obj ? cube_obj : "";
        a.earth_w = earth_w; a.earth_h = earth_h; a.earth_rgb = earth_rgb;
        Handle* h = new Handle();
        h->bw = build_world(name, a);
        return h;
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}

// Build a world from flat arrays (random / adversarial test scenes authored in Python).
// geom: 24 floats per item — sphere: cx cy cz r | rect: plane dist r0lo r0hi r1lo r1hi | tri: 9 pos, 9 nrm, 6 uv.
// mats: (kind, texture, param-as-float-bits) u32 triples.  texs: (kind, image, r, g, b, inv_frequency as float bits, even, odd) u32 x8.
// wrap (optional, may be NULL): 10 floats per item — flags (bit0 Transformation, bit1 Animation, bit2 Volume), offset xyz, y_sine, y_cosine, velocity xyz, neg_inv_density.
// perlins: n_perlins tables of 256 gradient vectors (xyz) + 3 x 256 permutation entries, all as floats / u32 in one block of 256*3 + 768 words each.
// cam: position3 ulc3 unit_right3 unit_up3 scaled_right3 scaled_up3 lens_radius t0 t1 pace2 (23 floats).
// bg: sky-flag r g b.  light: has plane dist r0lo r0hi r1lo r1hi (7 floats, has==0 -> none).
void* orc_world_create_flat(uint32_t n_items, const uint32_t* kinds, const uint32_t* item_mat, const float* geom,
                            uint32_t n_mats, const uint32_t* mats, uint32_t n_texs, const uint32_t* texs,
                            uint32_t n_images, const uint32_t* image_wh, const uint8_t* const* image_rgb,
                            const float* cam, const float* bg, const float* light, const float* wrap, uint32_t n_perlins,
                            const uint32_t* perlin_words) {
    try {
        if (n_items == 0) throw std::runtime_error("empty scene (reference recurses forever)");
        Handle* h = new Handle();
        h->bw.reset(new BuiltWorld());
        h->bw->wb.reset(new WorldBuilder());
        WorldBuilder& wb = *h->bw->wb;
        for (uint32_t i = 0; i < n_images; ++i) {
            Image im; im.width = image_wh[2 * i]; im.height = image_wh[2 * i + 1];
            im.rgb.assign(image_rgb[i], image_rgb[i] + (size_t)im.width * im.height * 3);
            wb.images.push_back(std::move(im));
        }
        for (uint32_t i = 0; i < n_perlins; ++i) {
            const uint32_t* pw = perlin_words + (size_t)i * 1536;
            Perlin pn;
            pn.bits = 8;
            for (int k = 0; k < 256; ++k) pn.ranvec.push_back(Dir3(bits_to_f32(pw[3 * k]), bits_to_f32(pw[3 * k + 1]), bits_to_f32(pw[3 * k + 2])));
            pn.perm_x.assign(pw + 768, pw + 1024); pn.perm_y.assign(pw + 1024, pw + 1280); pn.perm_z.assign(pw + 1280, pw + 1536);
            wb.perlins.push_back(pn);
        }
        std::vector<const Texture*> tptr;
        for (uint32_t i = 0; i < n_texs; ++i) {
            Texture t;
            t.kind = (Texture::Kind)texs[8 * i];
            if (t.kind == Texture::Marble) { t.noise = &wb.perlins.at(texs[8 * i + 1]); t.scale = bits_to_f32(texs[8 * i + 5]); }
            t.color = Color(bits_to_f32(texs[8 * i + 2]), bits_to_f32(texs[8 * i + 3]), bits_to_f32(texs[8 * i + 4]));
            if (t.kind == Texture::Image_) t.image = &wb.images[texs[8 * i + 1]];
            t.inv_frequency = bits_to_f32(texs[8 * i + 5]);
            wb.textures.push_back(t);
            tptr.push_back(&wb.textures.back());
        }
        for (uint32_t i = 0; i < n_texs; ++i)
            if (texs[8 * i] == Texture::Checker) {
                wb.textures[i].even = tptr.at(texs[8 * i + 6]);
                wb.textures[i].odd = tptr.at(texs[8 * i + 7]);
            }
        std::vector<const Material*> mptr;
        for (uint32_t i = 0; i < n_mats; ++i) {
            uint32_t k = mats[3 * i];
            const Texture* tx = (k == Material::Dielectric) ? nullptr : tptr.at(mats[3 * i + 1]);
            mptr.push_back(wb.mk((Material::Kind)k, tx, bits_to_f32(mats[3 * i + 2])));
        }
        World& w = h->bw->world;
        for (uint32_t i = 0; i < n_items; ++i) {
            const float* g = geom + 24 * (size_t)i;
            Geometry geo;
            if (kinds[i] == 0) {
                geo.kind = Geometry::Sphere;
                geo.sphere.center = Point3(g[0], g[1], g[2]); geo.sphere.radius = g[3];
            } else if (kinds[i] == 1) {
                geo.kind = Geometry::Rect;
                geo.rect.rect_plane = (int)g[0]; geo.rect.dist = g[1];
                geo.rect.r0[0] = g[2]; geo.rect.r0[1] = g[3]; geo.rect.r1[0] = g[4]; geo.rect.r1[1] = g[5];
            } else if (kinds[i] == 3) {
                geo.kind = Geometry::AxisAlignedBox;
                geo.box.min = Point3(g[0], g[1], g[2]); geo.box.max = Point3(g[3], g[4], g[5]);
            } else {
                geo.kind = Geometry::Triangle;
                for (int v = 0; v < 3; ++v) {
                    geo.tri.positions[v] = Point3(g[3 * v], g[3 * v + 1], g[3 * v + 2]);
                    geo.tri.normals[v] = Dir3(g[9 + 3 * v], g[10 + 3 * v], g[11 + 3 * v]);
                    geo.tri.texture_coords[v] = Vec2f(g[18 + 2 * v], g[19 + 2 * v]);
                }
            }
            SceneItem it;
            it.geo = geo;
            it.material = mptr.at(item_mat[i]);
            if (wrap) {
                const float* wv = wrap + 10 * (size_t)i;
                uint32_t fl = (uint32_t)wv[0];
                it.has_transform = (fl & 1u) != 0;
                it.transform.offset = Dir3(wv[1], wv[2], wv[3]); it.transform.y_sine = wv[4]; it.transform.y_cosine = wv[5];
                it.has_animation = (fl & 2u) != 0;
                it.velocity = Dir3(wv[6], wv[7], wv[8]);
                it.is_volume = (fl & 4u) != 0;
                it.neg_inv_density = wv[9];
            }
            w.scene.items.push_back(it);
        }
        w.scene.build();
        Camera& c = w.camera;
        c.position = Point3(cam[0], cam[1], cam[2]);
        c.upper_left_corner = Dir3(cam[3], cam[4], cam[5]);
        c.unit_right = Dir3(cam[6], cam[7], cam[8]);
        c.unit_up = Dir3(cam[9], cam[10], cam[11]);
        c.scaled_right = Dir3(cam[12], cam[13], cam[14]);
        c.scaled_up = Dir3(cam[15], cam[16], cam[17]);
        c.lens_radius = cam[18]; c.t0 = cam[19]; c.t1 = cam[20];
        c.shutter_pace = Vec2f(cam[21], cam[22]);
        w.background.sky = bg[0] != 0.0f;
        w.background.color = Color(bg[1], bg[2], bg[3]);
        w.has_light = light && light[0] != 0.0f;
        if (w.has_light) {
            w.light.rect_plane = (int)light[1]; w.light.dist = light[2];
            w.light.r0[0] = light[3]; w.light.r0[1] = light[4]; w.light.r1[0] = light[5]; w.light.r1[1] = light[6];
        }
        return h;
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}

void orc_world_destroy(void* hp) { delete (Handle*)hp; }

// info: n_items, n_nodes, root (encoded), depth, default_w, default_h, n_materials, n_textures, n_images, has_light
void orc_world_info(void* hp, uint64_t* info) {
    Handle* h = (Handle*)hp;
    const World& w = h->bw->world;
    info[0] = w.scene.items.size();
    info[1] = w.scene.nodes.size();
    info[2] = enc_child(w.scene.initial_index, w.scene.items.size());
    info[3] = w.scene.depth;
    info[4] = (uint64_t)h->bw->default_width;
    info[5] = (uint64_t)h->bw->default_height;
    info[6] = h->bw->wb->materials.size();
    info[7] = h->bw->wb->textures.size();
    info[8] = h->bw->wb->images.size();
    info[9] = w.has_light ? 1 : 0;
    info[10] = h->bw->wb->perlins.size();
}
// 1536 words per table: 256 x (x, y, z) float bits, then perm_x, perm_y, perm_z (256 u32 each)
void orc_world_perlin(void* hp, uint32_t index, uint32_t* words) {
    const Perlin& p = ((Handle*)hp)->bw->wb->perlins.at(index);
    for (int k = 0; k < 256; ++k) for (int c = 0; c < 3; ++c) words[3 * k + c] = f32_to_bits(p.ranvec[k].e[c]);
    for (int k = 0; k < 256; ++k) { words[768 + k] = p.perm_x[k]; words[1024 + k] = p.perm_y[k]; words[1280 + k] = p.perm_z[k]; }
}

// Flat export; array layouts as in orc_world_create_flat.  nodes_box: 6 floats (min, max); nodes_lra: left, right, axis.
void orc_world_export(void* hp, uint32_t* kinds, uint32_t* item_mat, float* geom, float* nodes_box, uint32_t* nodes_lra,
                      uint32_t* mats, uint32_t* texs, uint32_t* image_wh, float* cam, float* bg, float* light, float* wrap) {
    Handle* h = (Handle*)hp;
    const World& w = h->bw->world;
    const WorldBuilder& wb = *h->bw->wb;
    auto mat_index = [&](const Material* m) -> uint32_t {
        for (size_t i = 0; i < wb.materials.size(); ++i) if (&wb.materials[i] == m) return (uint32_t)i;
        return 0xFFFFFFFFu;
    };
    auto tex_index = [&](const Texture* t) -> uint32_t {
        for (size_t i = 0; i < wb.textures.size(); ++i) if (&wb.textures[i] == t) return (uint32_t)i;
        return 0xFFFFFFFFu;
    };
    auto img_index = [&](const Image* t) -> uint32_t {
        for (size_t i = 0; i < wb.images.size(); ++i) if (&wb.images[i] == t) return (uint32_t)i;
        return 0xFFFFFFFFu;
    };
    // material pointer -> index map (linear scan is too slow for 1M items: cache last)
    const Material* last_m = nullptr; uint32_t last_i = 0;
    for (size_t i = 0; i < w.scene.items.size(); ++i) {
        const SceneItem& it = w.scene.items[i];
        float* g = geom + 24 * i;
        std::memset(g, 0, 24 * sizeof(float));
        kinds[i] = (uint32_t)it.geo.kind;
        if (it.material != last_m) { last_m = it.material; last_i = mat_index(it.material); }
        item_mat[i] = last_i;
        if (wrap) {
            float* wv = wrap + 10 * i;
            wv[0] = (float)((it.has_transform ? 1 : 0) | (it.has_animation ? 2 : 0) | (it.is_volume ? 4 : 0));
            for (int k = 0; k < 3; ++k) { wv[1 + k] = it.transform.offset.e[k]; wv[6 + k] = it.velocity.e[k]; }
            wv[4] = it.transform.y_sine; wv[5] = it.transform.y_cosine;
            wv[9] = it.neg_inv_density;
        }
        if (it.geo.kind == Geometry::Sphere) {
            g[0] = it.geo.sphere.center.e[0]; g[1] = it.geo.sphere.center.e[1]; g[2] = it.geo.sphere.center.e[2];
            g[3] = it.geo.sphere.radius;
        } else if (it.geo.kind == Geometry::Rect) {
            g[0] = (float)it.geo.rect.rect_plane; g[1] = it.geo.rect.dist;
            g[2] = it.geo.rect.r0[0]; g[3] = it.geo.rect.r0[1]; g[4] = it.geo.rect.r1[0]; g[5] = it.geo.rect.r1[1];
        } else if (it.geo.kind == Geometry::AxisAlignedBox) {
            for (int k = 0; k < 3; ++k) { g[k] = it.geo.box.min.e[k]; g[3 + k] = it.geo.box.max.e[k]; }
        } else {
            for (int v = 0; v < 3; ++v) {
                for (int k = 0; k < 3; ++k) { g[3 * v + k] = it.geo.tri.positions[v].e[k]; g[9 + 3 * v + k] = it.geo.tri.normals[v].e[k]; }
                g[18 + 2 * v] = it.geo.tri.texture_coords[v].x; g[19 + 2 * v] = it.geo.tri.texture_coords[v].y;
            }
        }
    }
    for (size_t i = 0; i < w.scene.nodes.size(); ++i) {
        const BvhNode& n = w.scene.nodes[i];
        for (int k = 0; k < 3; ++k) { nodes_box[6 * i + k] = n.aabb.min.e[k]; nodes_box[6 * i + 3 + k] = n.aabb.max.e[k]; }
        nodes_lra[3 * i] = enc_child(n.left, w.scene.items.size());
        nodes_lra[3 * i + 1] = enc_child(n.right, w.scene.items.size());
        nodes_lra[3 * i + 2] = (uint32_t)n.axis_id;
    }
    for (size_t i = 0; i < wb.materials.size(); ++i) {
        mats[3 * i] = (uint32_t)wb.materials[i].kind;
        mats[3 * i + 1] = wb.materials[i].tex ? tex_index(wb.materials[i].tex) : 0xFFFFFFFFu;
        mats[3 * i + 2] = f32_to_bits(wb.materials[i].param);
    }
    for (size_t i = 0; i < wb.textures.size(); ++i) {
        texs[8 * i] = (uint32_t)wb.textures[i].kind;
        texs[8 * i + 1] = wb.textures[i].image ? img_index(wb.textures[i].image) : 0xFFFFFFFFu;
        for (int k = 0; k < 3; ++k) texs[8 * i + 2 + k] = f32_to_bits(wb.textures[i].color.e[k]);
        texs[8 * i + 5] = f32_to_bits(wb.textures[i].kind == Texture::Marble ? wb.textures[i].scale : wb.textures[i].inv_frequency);
        if (wb.textures[i].kind == Texture::Marble)
            for (size_t k = 0; k < wb.perlins.size(); ++k) if (&wb.perlins[k] == wb.textures[i].noise) texs[8 * i + 1] = (uint32_t)k;
        texs[8 * i + 6] = wb.textures[i].even ? tex_index(wb.textures[i].even) : 0xFFFFFFFFu;
        texs[8 * i + 7] = wb.textures[i].odd ? tex_index(wb.textures[i].odd) : 0xFFFFFFFFu;
    }
    for (size_t i = 0; i < wb.images.size(); ++i) { image_wh[2 * i] = wb.images[i].width; image_wh[2 * i + 1] = wb.images[i].height; }
    const Camera& c = w.camera;
    const Dir3* vs[6] = {&c.position, &c.upper_left_corner, &c.unit_right, &c.unit_up, &c.scaled_right, &c.scaled_up};
    for (int v = 0; v < 6; ++v) for (int k = 0; k < 3; ++k) cam[3 * v + k] = vs[v]->e[k];
    cam[18] = c.lens_radius; cam[19] = c.t0; cam[20] = c.t1; cam[21] = c.shutter_pace.x; cam[22] = c.shutter_pace.y;
    bg[0] = w.background.sky ? 1.0f : 0.0f;
    for (int k = 0; k < 3; ++k) bg[1 + k] = w.background.color.e[k];
    light[0] = w.has_light ? 1.0f : 0.0f;
    light[1] = (float)w.light.rect_plane; light[2] = w.light.dist;
    light[3] = w.light.r0[0]; light[4] = w.light.r0[1]; light[5] = w.light.r1[0]; light[6] = w.light.r1[1];
}
const uint8_t* orc_world_image(void* hp, uint32_t i) { return ((Handle*)hp)->bw->wb->images[i].rgb.data(); }

// Scene::hit on caller rays.  hit_id = flattened item index, 0xFFFFFFFF = miss.  counters (optional): node_tests, prim_tests.
static float g_trace_time = 0.0f;
void orc_set_trace_time(float t) { g_trace_time = t; }  // Ray::time of the rays handed to orc_trace (Animation leaves)
void orc_trace(void* hp, uint64_t n, const float* org, const float* dir, float t_min, float t_max, uint32_t* hit_id,
               float* t, float* normal, uint8_t* front_face, float* uv, uint64_t* counters) {
    const World& w = ((Handle*)hp)->bw->world;
    Counters cnt;
    for (uint64_t i = 0; i < n; ++i) {
        Ray r{Point3(org[3 * i], org[3 * i + 1], org[3 * i + 2]), Dir3(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]), g_trace_time};
        Range tr{t_min, t_max};
        HitInteraction hi;
        // Scene::hit takes an rng (volumes draw inside the traversal): rt_trace keys it by the ray index (REPLAY, seed 0)
        TRng rng = TRng::make_replay(0);
        rng.set_path((uint32_t)i, (uint32_t)(i >> 32));
        rng.begin_block(0);
        if (w.scene.hit(r, tr, rng, &hi, &cnt)) {
            hit_id[i] = hi.item; t[i] = hi.t;
            for (int k = 0; k < 3; ++k) normal[3 * i + k] = hi.normal.e[k];
            front_face[i] = hi.front_face ? 1 : 0;
            uv[2 * i] = hi.uv.x; uv[2 * i + 1] = hi.uv.y;
        } else {
            hit_id[i] = 0xFFFFFFFFu; t[i] = INF_F;
            for (int k = 0; k < 3; ++k) normal[3 * i + k] = 0.0f;
            front_face[i] = 0; uv[2 * i] = 0.0f; uv[2 * i + 1] = 0.0f;
        }
    }
    if (counters) { counters[0] = cnt.node_tests; counters[1] = cnt.prim_tests; }
}

// Pixel-grid primary rays exactly as render() generates them with zero jitter and the lens disabled
// (Size2i::iterf, src/lib/size2i.rs:38-56; Camera::ray, src/lib/camera.rs:175-200).
void orc_primary_rays(void* hp, int width, int height, float* org, float* dir) {
    const World& w = ((Handle*)hp)->bw->world;
    Camera cam = w.camera;
    cam.lens_radius = 0.0f;
    cam.t1 = cam.t0;
    TRng rng = TRng::from_u64(0);
    float sx = 1.0f / (float)(width - 1), sy = 1.0f / (float)(height - 1);
    for (int y = 0; y < height; ++y)
        for (int x = 0; x < width; ++x) {
            Ray r = cam.ray(rng, Vec2f((float)x * sx, (float)y * sy));
            size_t i = (size_t)y * width + x;
            for (int k = 0; k < 3; ++k) { org[3 * i + k] = r.origin.e[k]; dir[3 * i + k] = r.direction.e[k]; }
        }
}

// render()  — src/lib/rendering.rs:121-220.  mode: 0 = FAITHFUL (Xoroshiro, rejection samplers), 1 = REPLAY
// (Philox (seed, pixel, sample, block) + inversion samplers; global sample index = k * n_tasks + task id).
// render_mode: 0 = Default, 1 = Normals.  stats: [0] ray segments, [1] node tests, [2] prim tests.  Returns seconds
// (timing brackets = the reference's: thread spawn .. join, excluding merge).
// lane_first / lane_stride select the subset of worker tasks this call renders (multi-GPU split of the same
// function); apply_scale = 0 returns the un-normalised plane sum (before the 1/len multiplier of merge_planes).
static double render_impl(void* hp, int width, int height, uint64_t spp, int max_depth, uint64_t thread_count, uint64_t seed,
                          int mode, int render_mode, uint32_t lane_first, uint32_t lane_stride, int apply_scale,
                          float* out_rgb, uint64_t* stats) {
    const World& world = ((Handle*)hp)->bw->world;
    const size_t npix = (size_t)width * height;
    UniformF32 jx(0.0f, 1.0f / (float)(width - 1)), jy(0.0f, 1.0f / (float)(height - 1));  // :132-138
    std::vector<std::pair<size_t, size_t>> all_tasks = split_work_tasks(spp, thread_count);
    const size_t n_tasks = all_tasks.size();
    std::vector<std::pair<size_t, size_t>> tasks;
    for (size_t i = lane_first; i < n_tasks; i += lane_stride) tasks.push_back(all_tasks[i]);
    std::vector<std::vector<Color>> planes(tasks.size());
    std::vector<Counters> cnts(tasks.size());
    auto t_start = std::chrono::steady_clock::now();
    std::vector<std::thread> threads;
    for (size_t ti = 0; ti < tasks.size(); ++ti) {
        threads.emplace_back([&, ti]() {
            const size_t id = tasks[ti].first, real_spp = tasks[ti].second;
            TRng rng = mode == 1 ? TRng::make_replay(seed) : TRng::from_u64(seed * 0x9E3779B97F4A7C15ull + id + 1);
            std::vector<Color>& plane = planes[ti];
            plane.resize(npix);
            Counters& cnt = cnts[ti];
            float sx = 1.0f / (float)(width - 1), sy = 1.0f / (float)(height - 1);  // size2i.rs:21-22
            for (int y = 0; y < height; ++y)
                for (int x = 0; x < width; ++x) {
                    Vec2f fpix((float)x * sx, (float)y * sy);
                    Color sum(0, 0, 0);
                    for (size_t k = 0; k < real_spp; ++k) {
                        if (mode == 1) { rng.set_path((uint32_t)((size_t)y * width + x), (uint32_t)(k * n_tasks + id)); rng.begin_block(0); }
                        float jxv = rng.uniform(jx);
                        float jyv = rng.uniform(jy);
                        Vec2f pix = fpix + Vec2f(jxv, jyv);
                        Ray ray = world.camera.ray(rng, pix);
                        Color c;
                        if (render_mode == 0) {
                            c = ray_color(ray, world, rng, max_depth, &cnt);
                        } else {  // RenderMode::Normals :108-116
                            rng.begin_block(1);  // REPLAY bookkeeping: the first segment draws from block 1, as in ray_color
                            HitInteraction hi;
                            Range tr{0.001f, INF_F};
                            if (world.scene.hit(ray, tr, rng, &hi, &cnt)) c = (hi.normal + V3(1.0f, 1.0f, 1.0f)) * 0.5f;
                            else c = world.background.sample(ray);
                        }
                        sum = sum + c;
                    }
                    plane[(size_t)y * width + x] = sum / (float)real_spp;
                }
        });
    }
    for (auto& th : threads) th.join();
    double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
    std::vector<Color> pixels(npix);
    if (!planes.empty()) {
        // merge_planes :239-252 — the LAST plane is popped first; a rank that does not own the globally last task sums ascending
        const bool owns_last = tasks.back().first == n_tasks - 1;
        if (owns_last) {
            pixels = std::move(planes.back());
            planes.pop_back();
        } else {
            pixels = std::move(planes.front());
            planes.erase(planes.begin());
        }
        for (const auto& plane : planes)
            for (size_t i = 0; i < npix; ++i) pixels[i] = pixels[i] + plane[i];
        if (apply_scale) {
            float multiplier = 1.0f / (float)n_tasks;
            for (size_t i = 0; i < npix; ++i) pixels[i] = pixels[i] * multiplier;
        }
    }
    for (size_t i = 0; i < npix; ++i) for (int k = 0; k < 3; ++k) out_rgb[3 * i + k] = pixels[i].e[k];
    if (stats) {
        stats[0] = stats[1] = stats[2] = 0;
        for (const Counters& c : cnts) { stats[0] += c.rays; stats[1] += c.node_tests; stats[2] += c.prim_tests; }
    }
    return secs;
}
double orc_render(void* hp, int width, int height, uint64_t spp, int max_depth, uint64_t thread_count, uint64_t seed,
                  int mode, int render_mode, float* out_rgb, uint64_t* stats) {
    return render_impl(hp, width, height, spp, max_depth, thread_count, seed, mode, render_mode, 0, 1, 1, out_rgb, stats);
}
double orc_render_planes(void* hp, int width, int height, uint64_t spp, int max_depth, uint64_t thread_count, uint64_t seed,
                         int mode, int render_mode, uint32_t lane_first, uint32_t lane_stride, float* out_rgb, uint64_t* stats) {
    return render_impl(hp, width, height, spp, max_depth, thread_count, seed, mode, render_mode, lane_first, lane_stride, 0, out_rgb, stats);
}

// Color::to_rgb8_gamma2 — src/lib/color.rs:43-48 (caller side, src/app/main.rs:53-56)
void orc_to_rgb8_gamma2(const float* rgb, uint64_t n_pixels, uint8_t* out) {
    for (uint64_t i = 0; i < n_pixels; ++i) to_rgb8_gamma2(Color(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]), out + 3 * i);
}

// Unit-level probes for the known-answer tests.
int orc_aabb_hit_cond(const float* box6, const float* org, const float* dir, float t0, float t1) {
    Aabb b; b.min = Point3(box6[0], box6[1], box6[2]); b.max = Point3(box6[3], box6[4], box6[5]);
    Ray r{Point3(org[0], org[1], org[2]), Dir3(dir[0], dir[1], dir[2]), 0.0f};
    return aabb_hit_cond(b, r, Range{t0, t1}) ? 1 : 0;
}
void orc_obj_load(const char* path, uint64_t* n_tris, float* out24, uint64_t cap) {
    try {
        std::vector<TriangleGeometry> t = load_obj_file(path);
        *n_tris = t.size();
        for (size_t i = 0; i < t.size() && i < cap; ++i) {
            float* g = out24 + 24 * i;
            for (int v = 0; v < 3; ++v) {
                for (int k = 0; k < 3; ++k) { g[3 * v + k] = t[i].positions[v].e[k]; g[9 + 3 * v + k] = t[i].normals[v].e[k]; }
                g[18 + 2 * v] = t[i].texture_coords[v].x; g[19 + 2 * v] = t[i].texture_coords[v].y;
            }
        }
    } catch (const std::exception& e) { g_err = e.what(); *n_tris = (uint64_t)-1; }
}
void orc_rng_kat(uint64_t* u64_out, float* f32_out) {  // first draws of the scene RNG (src/app/main.rs:24)
    TRng r = scene_rng();
    for (int i = 0; i < 4; ++i) u64_out[i] = r.next_u64();
    TRng r2 = scene_rng();
    for (int i = 0; i < 4; ++i) f32_out[i] = r2.gen_f32();
}
void orc_xoroshiro_from_seed(const uint8_t* seed16, uint64_t* out, int n) {  // Xoroshiro128PlusPlus::from_seed + next_u64
    TRng r = TRng::from_seed_bytes(seed16);
    for (int i = 0; i < n; ++i) out[i] = r.next_u64();
}
void orc_probe_math(float u, float* out3) {  // sincos_2pi / cbrt_unit of the REPLAY samplers
    sincos_2pi(u, &out3[0], &out3[1]);
    out3[2] = cbrt_unit(u);
}
float orc_probe_ln(float u) { return ln_unit(u); }  // free-flight ln of the REPLAY volume sampler
int orc_split_work_tasks(uint64_t spp, uint64_t threads, uint64_t* out_counts, uint64_t cap) {
    auto t = split_work_tasks(spp, threads);
    for (size_t i = 0; i < t.size() && i < cap; ++i) out_counts[i] = t[i].second;
    return (int)t.size();
}
void orc_philox(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t* out) {
    philox4x32_10(k0, k1, c0, c1, c2, c3, out);
}

}  // extern "C"
