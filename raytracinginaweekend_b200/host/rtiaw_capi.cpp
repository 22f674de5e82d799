// rtiaw_capi.cpp — C shim (include/rtiaw_host.h) over the C++ host mirror.
#include "rtiaw_host.h"

#include "rtiaw_app.hpp"

using namespace rtiaw;

struct RtwWorld {
    std::unique_ptr<World> world;
    Size2i default_size{0, 0};
};

namespace {
thread_local std::string g_err;
template <class F>
int guard(F&& f) {
    try {
        f();
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}
}  // namespace

extern "C" {

const char* rtw_last_error(void) { return g_err.c_str(); }

RtwWorld* rtw_world_create(const char* name, const char* suzanne, const char* cube, uint32_t ew, uint32_t eh, const uint8_t* ergb) {
    RtwWorld* w = nullptr;
    guard([&] {
        Assets a;
        a.suzanne_obj = suzanne ? suzanne : "";
        a.cube_obj = cube ? cube : "";
        a.earth_w = ew; a.earth_h = eh; a.earth_rgb8 = ergb;
        NamedWorld nw = create_world(name ? name : "", a);
        w = new RtwWorld{std::move(nw.world), nw.default_size};
    });
    return w;
}

RtwWorld* rtw_world_create_elements(uint32_t n, const uint32_t* kinds, const uint32_t* mat_of, const float* geom, uint32_t n_materials,
                                    const RtMaterial* materials, uint32_t n_textures, const RtTexture* textures, uint32_t n_images,
                                    const RtImage* images, const float* cam, const float* bg, const float* light, const float* wrap,
                                    uint32_t n_perlins, const RtPerlin* perlins) {
    RtwWorld* w = nullptr;
    guard([&] {
        auto tables = std::make_shared<MaterialTable>();
        for (uint32_t i = 0; i < n_images; ++i) {
            ImageData im;
            im.width = images[i].width; im.height = images[i].height;
            im.rgb8.assign(images[i].rgb8, images[i].rgb8 + (size_t)im.width * im.height * 3);
            tables->images.push_back(std::move(im));
        }
        for (uint32_t i = 0; i < n_perlins; ++i) {
            PerlinData pn;
            pn.ranvec.assign(perlins[i].ranvec, perlins[i].ranvec + 768);
            pn.perm_x.assign(perlins[i].perm_x, perlins[i].perm_x + 256);
            pn.perm_y.assign(perlins[i].perm_y, perlins[i].perm_y + 256);
            pn.perm_z.assign(perlins[i].perm_z, perlins[i].perm_z + 256);
            tables->perlins.push_back(std::move(pn));
        }
        tables->textures.assign(textures, textures + n_textures);
        tables->materials.assign(materials, materials + n_materials);
        std::vector<SurfaceGeometry> elems(n);
        for (uint32_t i = 0; i < n; ++i) {
            const float* g = geom + 24 * (size_t)i;
            Geometry geo;
            if (kinds[i] == RT_GEO_SPHERE) {
                geo = Geometry::Sphere({{g[0], g[1], g[2]}, g[3]});
            } else if (kinds[i] == RT_GEO_RECT) {
                RectGeometry r;
                r.rect_plane = (RectPlane)(uint32_t)g[0]; r.dist = g[1];
                r.r0[0] = g[2]; r.r0[1] = g[3]; r.r1[0] = g[4]; r.r1[1] = g[5];
                geo = Geometry::Rect(r);
            } else if (kinds[i] == RT_GEO_TRIANGLE) {
                TriangleGeometry t;
                for (int k = 0; k < 3; ++k) {
                    t.positions[k] = {g[3 * k], g[3 * k + 1], g[3 * k + 2]};
                    t.normals[k] = {g[9 + 3 * k], g[10 + 3 * k], g[11 + 3 * k]};
                    t.texture_coords[k] = {g[18 + 2 * k], g[19 + 2 * k]};
                }
                geo = Geometry::Triangle(t);
            } else if (kinds[i] == RT_GEO_BOX) {
                geo = Geometry::AxisAlignedBox(Aabb{{g[0], g[1], g[2]}, {g[3], g[4], g[5]}});
            } else {
                throw std::invalid_argument("geometry kind outside the hot-path scope");
            }
            SurfaceGeometry e;
            e.geometry = geo;
            e.material = mat_of[i];
            if (wrap) {  // flags, offset xyz, y_sine, y_cosine, velocity xyz
                const float* wv = wrap + 10 * (size_t)i;
                const uint32_t fl = (uint32_t)wv[0];
                e.has_transformation = (fl & RT_WRAP_TRANSFORMATION) != 0;
                e.transformation.offset = {wv[1], wv[2], wv[3]};
                e.transformation.y_sine = wv[4]; e.transformation.y_cosine = wv[5];
                e.has_animation = (fl & RT_WRAP_ANIMATION) != 0;
                e.velocity = {wv[6], wv[7], wv[8]};
                e.is_volume = (fl & RT_WRAP_VOLUME) != 0;
                e.neg_inv_density = wv[9];
            }
            elems[i] = e;
        }
        auto world = std::make_unique<World>();
        world->tables = tables;
        world->hittable = std::make_shared<Scene>(std::move(elems));
        Camera& c = world->camera;
        c.position = {cam[0], cam[1], cam[2]};
        c.upper_left_corner = {cam[3], cam[4], cam[5]};
        c.unit_right = {cam[6], cam[7], cam[8]};
        c.unit_up = {cam[9], cam[10], cam[11]};
        c.scaled_right = {cam[12], cam[13], cam[14]};
        c.scaled_up = {cam[15], cam[16], cam[17]};
        c.lens_radius = cam[18];
        c.time_interval = {cam[19], cam[20]};
        c.shutter_pace = {cam[21], cam[22]};
        world->background = bg[0] != 0.0f ? BackgroundColor::Sky() : BackgroundColor::Solid({bg[1], bg[2], bg[3]});
        if (bg[0] != 0.0f) world->background.color = {bg[1], bg[2], bg[3]};
        if (light && light[0] != 0.0f) {
            world->has_scattering_distribution_provider = true;
            RectGeometry& r = world->scattering_distribution_provider.rect;
            r.rect_plane = (RectPlane)(uint32_t)light[1]; r.dist = light[2];
            r.r0[0] = light[3]; r.r0[1] = light[4]; r.r1[0] = light[5]; r.r1[1] = light[6];
        }
        w = new RtwWorld{std::move(world), Size2i{0, 0}};
    });
    return w;
}

void rtw_world_destroy(RtwWorld* w) { delete w; }

int rtw_world_info(const RtwWorld* w, uint64_t* info) {
    return guard([&] {
        const World& W = *w->world;
        info[0] = W.hittable->elements.size();
        info[1] = W.hittable->bvh.nodes.size();
        info[2] = W.hittable->bvh.initial_index;
        info[3] = W.hittable->bvh.depth;
        info[4] = (uint64_t)w->default_size.width;
        info[5] = (uint64_t)w->default_size.height;
        info[6] = W.tables->materials.size();
        info[7] = W.tables->textures.size();
        info[8] = W.tables->images.size();
        info[9] = W.has_scattering_distribution_provider ? 1 : 0;
        info[10] = W.tables->perlins.size();
    });
}

int rtw_world_export(const RtwWorld* w, RtItem* items, RtBvhNode* nodes, RtMaterial* materials, RtTexture* textures,
                     uint32_t* image_wh, float* cam, float* bg, float* light) {
    return guard([&] {
        const World& W = *w->world;
        World::ItemArray it;
        std::vector<RtImage> im;
        std::vector<RtPerlin> pn;
        RtSceneDesc d = W.describe(it, im, pn);
        std::memcpy(items, d.items, sizeof(RtItem) * d.n_items);
        if (d.n_nodes) std::memcpy(nodes, d.nodes, sizeof(RtBvhNode) * d.n_nodes);
        std::memcpy(materials, d.materials, sizeof(RtMaterial) * d.n_materials);
        if (d.n_textures) std::memcpy(textures, d.textures, sizeof(RtTexture) * d.n_textures);
        for (uint32_t i = 0; i < d.n_images; ++i) { image_wh[2 * i] = d.images[i].width; image_wh[2 * i + 1] = d.images[i].height; }
        std::memcpy(cam, &d.camera, sizeof(float) * 23);
        bg[0] = d.background_is_sky ? 1.0f : 0.0f;
        bg[1] = d.background_rgb[0]; bg[2] = d.background_rgb[1]; bg[3] = d.background_rgb[2];
        light[0] = d.has_light ? 1.0f : 0.0f;
        light[1] = (float)d.light_plane; light[2] = d.light_dist;
        light[3] = d.light_r0[0]; light[4] = d.light_r0[1]; light[5] = d.light_r1[0]; light[6] = d.light_r1[1];
    });
}

int rtw_world_perlin(const RtwWorld* w, uint32_t index, float* ranvec768, uint32_t* perm768) {
    return guard([&] {
        const PerlinData& p = w->world->tables->perlins.at(index);
        std::memcpy(ranvec768, p.ranvec.data(), 768 * sizeof(float));
        std::memcpy(perm768, p.perm_x.data(), 256 * 4);
        std::memcpy(perm768 + 256, p.perm_y.data(), 256 * 4);
        std::memcpy(perm768 + 512, p.perm_z.data(), 256 * 4);
    });
}

const uint8_t* rtw_world_image(const RtwWorld* w, uint32_t index) { return w->world->tables->images.at(index).rgb8.data(); }

int rtw_world_device(RtwWorld* w, int device, RtScene** out_scene) {
    return guard([&] { *out_scene = w->world->device(device); });
}

int rtw_world_release_device(RtwWorld* w) {
    return guard([&] { w->world->release_device(); });
}

int rtw_world_upload_bytes(const RtwWorld* w, uint64_t* out_bytes) {
    return guard([&] {
        const World& W = *w->world;
        uint64_t b = W.hittable->elements.size() * sizeof(RtItem) + W.hittable->bvh.nodes.size() * sizeof(RtBvhNode) +
                     W.tables->materials.size() * sizeof(RtMaterial) + W.tables->textures.size() * sizeof(RtTexture) + sizeof(RtSceneDesc);
        for (const ImageData& im : W.tables->images) b += im.rgb8.size();
        *out_bytes = b;
    });
}

int rtw_render(RtwWorld* w, int32_t width, int32_t height, uint64_t thread_count, uint64_t spp, int32_t max_depth, uint32_t render_mode,
               uint64_t seed, int device, float* out_rgb, RtStats* out_stats) {
    return guard([&] {
        render_into(Size2i{width, height}, thread_count, spp, max_depth, *w->world, (RenderMode)render_mode, reinterpret_cast<Color*>(out_rgb), seed, out_stats, device);
    });
}

int rtw_hit(RtwWorld* w, uint64_t n, const float* org, const float* dirs, float t_min, float t_max, int device, uint32_t* hit_id,
            float* t, float* normal, uint8_t* front_face, float* uv) {
    return guard([&] {
        if (rt_trace(w->world->device(device), org, dirs, n, t_min, t_max, hit_id, t, normal, front_face, uv) != RT_OK) throw_rt("rt_trace");
    });
}

int rtw_load_obj(const char* path, uint64_t* n_tris, float* out24, uint64_t cap) {
    return guard([&] {
        std::vector<TriangleGeometry> tris = load_obj_file(path);
        *n_tris = tris.size();
        for (size_t i = 0; i < tris.size() && i < cap; ++i) {
            RtItem it;
            Geometry::Triangle(tris[i]).write(it);
            std::memcpy(out24 + 24 * i, it.geom, sizeof(it.geom));
        }
    });
}

}  // extern "C"
