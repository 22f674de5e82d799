// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_core.hpp header).  PARITY UNPINNED.
//
// Literal restatement of the reference's scene-authoring side that produces hot-path inputs:
//   src/app/obj_loader.rs (bug-compatible fan triangulation, Q4), src/app/worlds/world_builder.rs
//   (node tree, NodeRef::finish flattening), the in-scope worlds of src/app/worlds/demo_worlds.rs,
//   src/app/main.rs:24 (scene RNG seed).  Plus the five BASELINE.json bench configs (C1..C5),
//   which are defined in DESIGN.md ("Workload definitions") and restated here independently of the
//   product's own authoring code.
#pragma once
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>

#include "oracle_core.hpp"

namespace orc {

// ---------------------------------------------------------------- src/app/obj_loader.rs
struct ObjScanner {
    const std::string& s;
    size_t pos = 0;
    explicit ObjScanner(const std::string& str) : s(str) {}
    bool eof() const { return pos >= s.size(); }
    char peek() const { return s[pos]; }
    bool try_char(char c) {
        if (!eof() && s[pos] == c) { ++pos; return true; }
        return false;
    }
    void at_least_one_ws() {  // :107-112
        if (!try_char(' ')) throw std::runtime_error("obj: ExpectedAtLeastOnWhitespace");
        any_ws();
    }
    void any_ws() { while (try_char(' ')) {} }
    std::string digits() {  // :117-125
        size_t st = pos;
        while (!eof() && s[pos] >= '0' && s[pos] <= '9') ++pos;
        if (pos == st) throw std::runtime_error("obj: ExpectedDigits");
        return s.substr(st, pos - st);
    }
    size_t take_usize() { return (size_t)std::strtoull(digits().c_str(), nullptr, 10); }
    float take_f32() {  // :130-143: '-'? digits ('.' digits)?  then str::parse::<f32> (correctly rounded)
        size_t st = pos;
        try_char('-');
        digits();
        if (try_char('.')) digits();
        return std::strtof(s.substr(st, pos - st).c_str(), nullptr);
    }
};

struct ObjVertexRef {
    size_t pos;
    long uv, nor;  // -1 = None
};

static inline std::vector<TriangleGeometry> load_obj_mesh(std::istream& in) {  // :7-22, 214-270
    std::vector<Point3> positions;
    std::vector<Dir3> normals;
    std::vector<Vec2f> texture_coords;
    std::vector<TriangleGeometry> triangles;
    std::string line;
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        ObjScanner sc(line);
        if (sc.eof()) continue;
        char c = sc.peek();
        if (c == 'v') {
            ++sc.pos;
            if (sc.try_char('n')) {
                sc.at_least_one_ws(); float x = sc.take_f32();
                sc.at_least_one_ws(); float y = sc.take_f32();
                sc.at_least_one_ws(); float z = sc.take_f32();
                normals.push_back(Dir3(x, y, z));
            } else if (sc.try_char('t')) {
                sc.at_least_one_ws(); float x = sc.take_f32();
                sc.at_least_one_ws(); float y = sc.take_f32();
                texture_coords.push_back(Vec2f(x, y));
            } else {
                sc.at_least_one_ws(); float x = sc.take_f32();
                sc.at_least_one_ws(); float y = sc.take_f32();
                sc.at_least_one_ws(); float z = sc.take_f32();
                positions.push_back(Point3(x, y, z));
            }
        } else if (c == 'f') {
            ++sc.pos;
            sc.at_least_one_ws();
            std::vector<ObjVertexRef> ids;
            while (!sc.eof()) {  // take_vertex :188-211
                ObjVertexRef v{0, -1, -1};
                v.pos = sc.take_usize() - 1;
                sc.any_ws();
                if (sc.try_char('/')) {
                    sc.any_ws();
                    if (sc.try_char('/')) {
                        sc.any_ws();
                        v.nor = (long)sc.take_usize() - 1;
                    } else {
                        v.uv = (long)sc.take_usize() - 1;
                        sc.any_ws();
                        if (sc.try_char('/')) {
                            sc.any_ws();
                            v.nor = (long)sc.take_usize() - 1;
                        }
                    }
                }
                sc.any_ws();
                ids.push_back(v);
            }
            TriangleGeometry tri;  // all ORIGIN / ZERO  (:227-231)
            for (size_t i = 0; i < ids.size(); ++i) {
                if (ids[i].pos >= positions.size()) throw std::runtime_error("obj: PosIdOutOfRange");
                Point3 position = positions[ids[i].pos];
                Vec2f uv(0, 0);
                if (ids[i].uv >= 0) {
                    if ((size_t)ids[i].uv >= texture_coords.size()) throw std::runtime_error("obj: TexIdOutOfRange");
                    uv = texture_coords[ids[i].uv];
                }
                if (ids[i].nor < 0) throw std::runtime_error("obj: face vertex without normal (reference: todo!())");
                if ((size_t)ids[i].nor >= normals.size()) throw std::runtime_error("obj: NorIdOutOfRange");
                Dir3 normal = normals[ids[i].nor];
                if (i < 2) {
                    tri.positions[i] = position; tri.normals[i] = normal; tri.texture_coords[i] = uv;
                } else {
                    // Q4: slot 1 <- slot 2 BEFORE slot 2 is written (first triangle gets ORIGIN / zero normal)
                    tri.positions[1] = tri.positions[2]; tri.normals[1] = tri.normals[2];
                    tri.texture_coords[1] = tri.texture_coords[2];
                    tri.positions[2] = position; tri.normals[2] = normal; tri.texture_coords[2] = uv;
                    triangles.push_back(tri);
                }
            }
        }
        // '#', 'o', anything else: ignored (:166-169)
    }
    return triangles;
}
static inline std::vector<TriangleGeometry> load_obj_file(const std::string& path) {
    std::ifstream f(path);
    if (!f) throw std::runtime_error("cannot open " + path);
    return load_obj_mesh(f);
}

// ---------------------------------------------------------------- src/app/worlds/world_builder.rs
struct GeoEntry {
    Geometry geo;
    const Material* material;
    bool is_poi;
    float density = 1.0f;  // < 1: VolumeGeometry (world_builder.rs:305-309)
};
struct Node {  // :194-199
    std::vector<GeoEntry> geo;
    Transformation transformation;
    Dir3 moving_animation;
    std::vector<std::shared_ptr<Node>> children;
};
typedef std::shared_ptr<Node> NodeRef;

struct WorldBuilder {
    std::deque<Texture> textures;
    std::deque<Material> materials;
    std::deque<Image> images;
    std::deque<Perlin> perlins;

    const Texture* texture_solid(Color c) {
        Texture t; t.kind = Texture::Solid; t.color = c;
        textures.push_back(t);
        return &textures.back();
    }
    const Texture* texture_image(uint32_t w, uint32_t h, const uint8_t* rgb) {
        Image im; im.width = w; im.height = h; im.rgb.assign(rgb, rgb + (size_t)w * h * 3);
        images.push_back(std::move(im));
        Texture t; t.kind = Texture::Image_; t.image = &images.back();
        textures.push_back(t);
        return &textures.back();
    }
    const Texture* texture_checker(float inv_frequency, const Texture* even, const Texture* odd) {  // world_builder.rs:38-40
        Texture t; t.kind = Texture::Checker; t.inv_frequency = inv_frequency; t.even = even; t.odd = odd;
        textures.push_back(t);
        return &textures.back();
    }
    const Texture* texture_marble(float scale, TRng& rng) {  // world_builder.rs:32-37: Perlin::new(8, rng)
        perlins.push_back(Perlin::make(8, rng));
        Texture t; t.kind = Texture::Marble; t.scale = scale; t.noise = &perlins.back();
        textures.push_back(t);
        return &textures.back();
    }
    const Material* material_isotropic_solid(Color c) { return mk(Material::Isotropic, texture_solid(c), 0.0f); }
    const Material* mk(Material::Kind k, const Texture* tex, float p) {
        Material m; m.kind = k; m.tex = tex; m.param = p;
        materials.push_back(m);
        return &materials.back();
    }
    const Material* material_lambert_solid(Color c) { return mk(Material::Lambert, texture_solid(c), 0.0f); }
    const Material* material_lambert(const Texture* t) { return mk(Material::Lambert, t, 0.0f); }
    const Material* material_metal_solid(Color c, float fuzz) { return mk(Material::Metal, texture_solid(c), fuzz); }
    const Material* material_dielectric(float ior) { return mk(Material::Dielectric, nullptr, ior); }
    const Material* material_diffuse_light_solid(Color c) { return mk(Material::DiffuseLight, texture_solid(c), 0.0f); }

    static Geometry geo_rect(int plane, Point3 center, float s0, float s1) {  // :77-86
        int a0, a1, n;
        rect_axes(plane, &a0, &a1, &n);
        Geometry g; g.kind = Geometry::Rect;
        g.rect.rect_plane = plane;
        g.rect.dist = center.e[n];
        g.rect.r0[0] = center.e[a0] - s0 * 0.5f; g.rect.r0[1] = center.e[a0] + s0 * 0.5f;
        g.rect.r1[0] = center.e[a1] - s1 * 0.5f; g.rect.r1[1] = center.e[a1] + s1 * 0.5f;
        return g;
    }
    static Geometry geo_box(float width, float height, float depth) {  // :88-93
        Geometry g; g.kind = Geometry::AxisAlignedBox;
        g.box.min = Point3(0, 0, 0); g.box.max = Point3(width, height, depth);
        return g;
    }
    static NodeRef new_obj_box(float w, float h, float d, const Material* m) { return new_obj(geo_box(w, h, d), m); }
    static Geometry geo_sphere(float radius) {  // :94-99
        Geometry g; g.kind = Geometry::Sphere;
        g.sphere.center = Point3(0, 0, 0); g.sphere.radius = radius;
        return g;
    }
    static NodeRef new_group() { return std::make_shared<Node>(); }
    static NodeRef new_obj(const Geometry& g, const Material* m) {
        NodeRef n = std::make_shared<Node>();
        n->geo.push_back(GeoEntry{g, m, false});
        return n;
    }
    static NodeRef new_obj_sphere(float radius, const Material* m) { return new_obj(geo_sphere(radius), m); }
    static NodeRef new_obj_rect(int plane, Point3 c, float s0, float s1, const Material* m) { return new_obj(geo_rect(plane, c, s0, s1), m); }
    static NodeRef new_mesh(const std::vector<TriangleGeometry>& tris, const Material* m, float scale = 1.0f) {  // :172-191
        NodeRef n = std::make_shared<Node>();
        n->geo.reserve(tris.size());
        for (const TriangleGeometry& t : tris) {
            Geometry g; g.kind = Geometry::Triangle; g.tri = t;
            if (scale != 1.0f) for (int i = 0; i < 3; ++i) g.tri.positions[i] = g.tri.positions[i] * scale;
            n->geo.push_back(GeoEntry{g, m, false});
        }
        return n;
    }
};
static inline NodeRef translate(NodeRef n, Dir3 o) { n->transformation = n->transformation.translate(o); return n; }
static inline NodeRef rotate_up(NodeRef n, float deg) { n->transformation = n->transformation.rotate_around_up_deg(deg); return n; }
static inline NodeRef set_poi(NodeRef n) { for (auto& g : n->geo) g.is_poi = true; return n; }
static inline NodeRef set_density(NodeRef n, float density) {  // set_all_geo_densitity :254-263
    if (!(density >= 0.0f && density < 1.0f)) throw std::runtime_error("Invalid densitity");
    for (auto& g : n->geo) g.density = density;
    return n;
}
static inline NodeRef animate_moving(NodeRef n, Dir3 v) { n->moving_animation = n->moving_animation + v; return n; }

// NodeRef::finish / finish_internal  (:272-329): bake what can be baked, wrap the remainder as a run-time Transformation,
// wrap moving nodes as Animation, first POI rect WITHOUT a remainder becomes the light-sampling provider.
static inline void finish_internal(const Node& node, const Transformation& parent, std::vector<SceneItem>& result,
                                   bool* has_light, RectGeometry* light) {
    Transformation full = parent.then(node.transformation);
    for (const GeoEntry& ge : node.geo) {
        bool has_rest = false;
        Transformation rest;
        Geometry real = ge.geo.partial_apply_transformation(full, &has_rest, &rest);
        SceneItem it;
        it.geo = real;
        it.material = ge.material;
        it.has_transform = has_rest;
        it.transform = rest;
        it.has_animation = node.moving_animation != Dir3(0, 0, 0);
        it.velocity = node.moving_animation;
        it.is_volume = ge.density < 1.0f;
        if (it.is_volume) it.neg_inv_density = -1.0f / ge.density;  // VolumeGeometry::new :301-307
        result.push_back(it);
        if (ge.is_poi && !has_rest && !*has_light && real.kind == Geometry::Rect) { *has_light = true; *light = real.rect; }
    }
    for (const NodeRef& c : node.children) finish_internal(*c, full, result, has_light, light);
}

struct BuiltWorld {
    std::unique_ptr<WorldBuilder> wb;
    World world;
    int default_width = 0, default_height = 0;
};
static inline void finish(BuiltWorld& bw, const NodeRef& root, BackgroundColor bg, const Camera& cam) {
    bw.world.camera = cam;
    bw.world.background = bg;
    bw.world.has_light = false;
    bw.world.scene.items.clear();
    finish_internal(*root, Transformation(), bw.world.scene.items, &bw.world.has_light, &bw.world.light);
    bw.world.scene.build();
}

// ---------------------------------------------------------------- scene generators
struct SceneAssets {
    std::string suzanne_obj, cube_obj;
    uint32_t earth_w = 0, earth_h = 0;
    const uint8_t* earth_rgb = nullptr;
};
static inline TRng scene_rng() {  // src/app/main.rs:24
    uint8_t seed[16];
    for (int i = 0; i < 16; ++i) seed[i] = (uint8_t)(i + 1);
    return TRng::from_seed_bytes(seed);
}
static inline Color gen_color(TRng& rng) { float r = rng.gen_f32(); float g = rng.gen_f32(); float b = rng.gen_f32(); return Color(r, g, b); }
static inline Color gen_muted_color(TRng& rng) { Color a = gen_color(rng); Color b = gen_color(rng); return convolution(a, b); }

// demo_worlds.rs:395-463 body, parameterised by grid half-extent (reference: 11) so C5 can reuse it.
static inline void add_random_sphere_field(WorldBuilder& wb, TRng& rng, NodeRef scene, int half) {
    const Material* mat_ground = wb.material_lambert_solid(Color(0.5f, 0.5f, 0.5f));
    float ground_radius = 1000.0f;
    Dir3 ground_center(0.0f, -ground_radius, 0.0f);
    scene->children.push_back(translate(WorldBuilder::new_obj_sphere(ground_radius, mat_ground), Dir3(0.0f, 0.0f - ground_radius, 0.0f)));
    const Material* material_glass = wb.material_dielectric(1.5f);
    for (int a = -half; a <= half; ++a)
        for (int b = -half; b <= half; ++b) {
            float ox = rng.gen_f32() * 0.9f;
            float oz = rng.gen_f32() * 0.9f;
            Dir3 center = Dir3((float)a, 0.2f, (float)b) + Dir3(ox, 0.0f, oz);
            if (length(center - Dir3(4.0f, 0.2f, 0.0f)) > 0.9f) {
                const Material* material;
                float material_sample = rng.gen_f32();
                if (material_sample < 0.8f) material = wb.material_lambert_solid(gen_muted_color(rng));
                else if (material_sample < 0.95f) { Color c = gen_color(rng); float fz = rng.gen_range(0.0f, 0.5f); material = wb.material_metal_solid(c, fz); }
                else material = material_glass;
                float small_radius = 0.2f;
                Dir3 real_center = ground_center + with_length(center - ground_center, ground_radius + small_radius);
                scene->children.push_back(translate(WorldBuilder::new_obj_sphere(small_radius, material), real_center));
            }
        }
    scene->children.push_back(translate(WorldBuilder::new_obj_sphere(1.0f, material_glass), Dir3(0.0f, 1.0f, 0.0f)));
    scene->children.push_back(translate(WorldBuilder::new_obj_sphere(1.0f, wb.material_lambert_solid(Color(0.4f, 0.2f, 0.1f))), Dir3(-4.0f, 1.0f, 0.0f)));
    scene->children.push_back(translate(WorldBuilder::new_obj_sphere(1.0f, wb.material_metal_solid(Color(0.7f, 0.6f, 0.5f), 0.0f)), Dir3(4.0f, 1.0f, 0.0f)));
}

static inline Camera cam_lookat(float fov, float aspect, Point3 pos, Point3 at) {
    CameraBuilder cb;
    cb.vertical_fov(fov, aspect).set_position(pos).look_at(DIR_UP, at);
    return cb.build();
}
static inline BackgroundColor bg_sky() { BackgroundColor b; b.sky = true; return b; }
static inline BackgroundColor bg_solid(Color c) { BackgroundColor b; b.sky = false; b.color = c; return b; }

// create_cornell_box_node  (demo_worlds.rs:216-264)
static inline NodeRef cornell_box_node(WorldBuilder& wb, float light_size) {
    const Material* red = wb.material_lambert_solid(Color(0.65f, 0.05f, 0.05f));
    const Material* white = wb.material_lambert_solid(Color(0.73f, 0.73f, 0.73f));
    const Material* green = wb.material_lambert_solid(Color(0.12f, 0.45f, 0.15f));
    const Material* light = wb.material_diffuse_light_solid(Color(15.0f, 15.0f, 15.0f));
    float hsize = 278.0f;
    NodeRef g = WorldBuilder::new_group();
    g->children.push_back(WorldBuilder::new_obj_rect(RECT_YZ, Point3(0.0f, hsize, hsize), 2.0f * hsize, 2.0f * hsize, red));
    g->children.push_back(WorldBuilder::new_obj_rect(RECT_YZ, Point3(2.0f * hsize, hsize, hsize), 2.0f * hsize, 2.0f * hsize, green));
    g->children.push_back(WorldBuilder::new_obj_rect(RECT_XZ, Point3(hsize, 0.0f, hsize), 2.0f * hsize, 2.0f * hsize, white));
    g->children.push_back(WorldBuilder::new_obj_rect(RECT_XZ, Point3(hsize, 2.0f * hsize, hsize), 2.0f * hsize, 2.0f * hsize, white));
    g->children.push_back(WorldBuilder::new_obj_rect(RECT_XY, Point3(hsize, hsize, 2.0f * hsize), 2.0f * hsize, 2.0f * hsize, white));
    g->children.push_back(set_poi(WorldBuilder::new_obj_rect(RECT_XZ, Point3(hsize, 2.0f * hsize - 1.0f, hsize), light_size, light_size, light)));
    return g;
}

// name: demo:suzanne | demo:final_scene1 | demo:simple_plane | demo:earth_mapped | demo:defocus_blur |
//       config:c1 .. config:c5  (BASELINE.json configs; DESIGN.md "Workload definitions")
static inline std::unique_ptr<BuiltWorld> build_world(const std::string& name, const SceneAssets& assets) {
    std::unique_ptr<BuiltWorld> bw(new BuiltWorld());
    bw->wb.reset(new WorldBuilder());
    WorldBuilder& wb = *bw->wb;
    TRng rng = scene_rng();
    NodeRef scene = WorldBuilder::new_group();

    if (name == "demo:suzanne") {  // demo_worlds.rs:6-28
        Camera cam = cam_lookat(40.0f, 3.0f / 4.0f, Point3(0.0f, 2.0f, 10.0f), Point3(0, 0, 0));
        const Material* mat_ground = wb.material_lambert_solid(Color(0.4f, 0.4f, 0.4f));
        const Material* mat_monkey = wb.material_lambert_solid(Color(1.0f, 0.2f, 0.2f));
        scene->children.push_back(WorldBuilder::new_mesh(load_obj_file(assets.suzanne_obj), mat_monkey));
        scene->children.push_back(translate(WorldBuilder::new_obj_sphere(1000.1f, mat_ground), Dir3(0.0f, -1000.0f, 0.0f)));
        finish(*bw, scene, bg_sky(), cam);
        bw->default_width = 400; bw->default_height = 300;
    } else if (name == "demo:final_scene1" || name == "config:c1") {  // demo_worlds.rs:395-463
        float aspect = (name == "config:c1") ? (800.0f / 1200.0f) : (9.0f / 16.0f);
        CameraBuilder cb;
        cb.vertical_fov(60.0f, aspect).set_position(Point3(13.0f, 2.0f, 3.0f)).look_at(DIR_UP, Point3(0, 0, 0))
            .set_focus_distance(10.0f).set_aperture(0.1f);
        add_random_sphere_field(wb, rng, scene, 11);
        finish(*bw, scene, bg_sky(), cb.build());
        bw->default_width = 1200; bw->default_height = (name == "config:c1") ? 800 : 675;
    } else if (name == "demo:cornell_box") {  // demo_worlds.rs:266-296
        float hsize = 278.0f;
        Camera cam = cam_lookat(40.0f, 1.0f, Point3(hsize, hsize, -800.0f), Point3(hsize, hsize, 0.0f));
        const Material* white = wb.material_lambert_solid(Color(0.73f, 0.73f, 0.73f));
        scene->children.push_back(cornell_box_node(wb, 130.0f));
        scene->children.push_back(translate(rotate_up(WorldBuilder::new_obj_box(165.0f, 330.0f, 165.0f, white), 15.0f), Dir3(265.0f, 0.0f, 295.0f)));
        scene->children.push_back(translate(rotate_up(WorldBuilder::new_obj_box(165.0f, 165.0f, 165.0f, white), -18.0f), Dir3(130.0f, 0.0f, 65.0f)));
        finish(*bw, scene, bg_solid(Color(0, 0, 0)), cam);
        bw->default_width = 600; bw->default_height = 600;
    } else if (name == "demo:cornell_box_smoke") {  // demo_worlds.rs:179-214
        float hsize = 278.0f;
        Camera cam = cam_lookat(40.0f, 1.0f, Point3(hsize, hsize, -800.0f), Point3(hsize, hsize, 0.0f));
        const Material* smoke_black = wb.material_isotropic_solid(Color(0.0f, 0.0f, 0.0f));
        const Material* smoke_white = wb.material_isotropic_solid(Color(1.0f, 1.0f, 1.0f));
        scene->children.push_back(cornell_box_node(wb, 130.0f));
        scene->children.push_back(set_density(translate(rotate_up(WorldBuilder::new_obj_box(165.0f, 330.0f, 165.0f, smoke_white), 15.0f), Dir3(265.0f, 0.0f, 295.0f)), 0.01f));
        scene->children.push_back(set_density(translate(rotate_up(WorldBuilder::new_obj_box(165.0f, 165.0f, 165.0f, smoke_black), -18.0f), Dir3(130.0f, 0.0f, 65.0f)), 0.01f));
        finish(*bw, scene, bg_solid(Color(0, 0, 0)), cam);
        bw->default_width = 600; bw->default_height = 600;
    } else if (name == "demo:perlin_spheres") {  // demo_worlds.rs:154-177
        Camera cam = cam_lookat(40.0f, 3.0f / 4.0f, Point3(13.0f, 2.0f, 3.0f), Point3(0, 0, 0));
        const Material* mat_noise = wb.material_lambert(wb.texture_marble(4.0f, rng));
        scene->children.push_back(translate(WorldBuilder::new_obj_sphere(1000.0f, mat_noise), Dir3(0.0f, 0.0f - 1000.0f, 0.0f)));
        scene->children.push_back(translate(WorldBuilder::new_obj_sphere(2.0f, mat_noise), Dir3(0.0f, 2.0f, 0.0f)));
        finish(*bw, scene, bg_sky(), cam);
        bw->default_width = 400; bw->default_height = 300;
    } else if (name == "demo:final_scene2") {  // demo_worlds.rs:30-152
        CameraBuilder cb;
        cb.vertical_fov(40.0f, 1.0f).set_position(Point3(478.0f, 278.0f, -600.0f)).look_at(DIR_UP, Point3(278.0f, 278.0f, 0.0f)).motion_blur(0.0f, 1.0f);
        const Material* mat_ground = wb.material_lambert_solid(Color(0.48f, 0.83f, 0.53f));
        NodeRef ground = WorldBuilder::new_group();
        {
            const int boxes_per_side = 20;
            Point3 rmin(-1000.0f, 0.0f, -1000.0f), rmax(1000.0f, 101.0f, 1000.0f);
            float width = (rmax.e[0] - rmin.e[0]) / (float)boxes_per_side;
            float height = rmax.e[1] - rmin.e[1];
            float depth = (rmax.e[2] - rmin.e[2]) / (float)boxes_per_side;
            for (int i = 0; i < boxes_per_side; ++i)
                for (int j = 0; j < boxes_per_side; ++j) {
                    Dir3 mn = rmin + Dir3((float)i * width, 0.0f, (float)j * depth) - Point3(0, 0, 0);
                    ground->children.push_back(translate(WorldBuilder::new_obj_box(width, height * rng.gen_f32(), depth, mat_ground), mn));
                }
        }
        const Material* mat_light = wb.material_diffuse_light_solid(Color(7.0f, 7.0f, 7.0f));
        NodeRef light = WorldBuilder::new_obj_rect(RECT_XZ, Point3(273.0f, 554.0f, 279.0f), 300.0f, 300.0f, mat_light);
        const Material* mat_glass = wb.material_dielectric(1.5f);
        const Material* mat_metal = wb.material_metal_solid(Color(0.8f, 0.8f, 0.9f), 1.0f);
        const Texture* tex_marble = wb.texture_marble(0.1f, rng);
        const Material* mat_marble = wb.material_lambert(tex_marble);
        const Material* mat_earth = wb.material_lambert(wb.texture_image(assets.earth_w, assets.earth_h, assets.earth_rgb));
        const Material* mat_fog = wb.material_isotropic_solid(Color(0.2f, 0.4f, 0.9f));
        const Material* mat_moving = wb.material_lambert_solid(Color(0.7f, 0.3f, 0.1f));
        NodeRef floating = WorldBuilder::new_group();
        floating->children.push_back(translate(WorldBuilder::new_obj_sphere(70.0f, mat_glass), Dir3(360.0f, 150.0f, 145.0f)));
        floating->children.push_back(set_density(translate(WorldBuilder::new_obj_sphere(70.0f, mat_fog), Dir3(360.0f, 150.0f, 145.0f)), 0.2f));
        floating->children.push_back(translate(WorldBuilder::new_obj_sphere(50.0f, mat_glass), Dir3(260.0f, 150.0f, 45.0f)));
        floating->children.push_back(translate(WorldBuilder::new_obj_sphere(50.0f, mat_metal), Dir3(0.0f, 150.0f, 145.0f)));
        floating->children.push_back(translate(WorldBuilder::new_obj_sphere(80.0f, mat_marble), Dir3(220.0f, 280.0f, 300.0f)));
        floating->children.push_back(translate(WorldBuilder::new_obj_sphere(100.0f, mat_earth), Dir3(400.0f, 200.0f, 400.0f)));
        floating->children.push_back(animate_moving(translate(WorldBuilder::new_obj_sphere(50.0f, mat_moving), Dir3(400.0f, 400.0f, 200.0f)), Dir3(30.0f, 0.0f, 0.0f)));
        NodeRef sphere_cube = WorldBuilder::new_group();
        {
            const Material* material = wb.material_lambert_solid(Color(0.75f, 0.75f, 0.75f));
            Transformation rot = Transformation().rotate_around_up_deg(15.0f);
            Dir3 dir_1 = rot.apply_direction(Dir3(1, 0, 0)), dir_2 = rot.apply_direction(Dir3(0, 1, 0)), dir_3 = rot.apply_direction(Dir3(0, 0, -1));
            Dir3 origin(-100.0f, 270.0f, 395.0f);
            for (int i = 0; i < 1000; ++i) {
                float r0 = rng.gen_f32(), r1 = rng.gen_f32(), r2 = rng.gen_f32();
                Dir3 p = origin + 165.0f * (dir_1 * r0 + dir_2 * r1 + dir_3 * r2);
                sphere_cube->children.push_back(translate(WorldBuilder::new_obj_sphere(10.0f, material), p));
            }
        }
        NodeRef obj_fog = set_density(WorldBuilder::new_obj_sphere(5000.0f, wb.material_isotropic_solid(Color(1.0f, 1.0f, 1.0f))), 0.0001f);
        scene->children.push_back(obj_fog);
        scene->children.push_back(ground);
        scene->children.push_back(sphere_cube);
        scene->children.push_back(floating);
        scene->children.push_back(light);
        finish(*bw, scene, bg_solid(Color(0, 0, 0)), cb.build());
        bw->default_width = 600; bw->default_height = 600;
    } else if (name == "demo:moving_spheres") {  // demo_worlds.rs:351-393
        CameraBuilder cb;
        cb.vertical_fov(60.0f, 16.0f / 19.0f).set_position(Point3(0.0f, 2.0f, 10.0f)).look_at(DIR_UP, Point3(0.0f, 2.0f, 0.0f)).motion_blur(0.0f, 0.5f);
        const Texture* tex_black = wb.texture_solid(Color(0.0f, 0.0f, 0.0f));
        const Texture* tex_white = wb.texture_solid(Color(1.0f, 1.0f, 1.0f));
        const Texture* tex_ground = wb.texture_checker(10.0f, tex_black, tex_white);
        const Material* mat_ground = wb.material_lambert(tex_ground);
        const Material* mat_red = wb.material_lambert_solid(Color(0.6f, 0.2f, 0.2f));
        const Material* mat_blue = wb.material_lambert_solid(Color(0.2f, 0.2f, 0.6f));
        scene->children.push_back(translate(WorldBuilder::new_obj_sphere(100.0f, mat_ground), Dir3(0.0f, -100.0f, 0.0f)));
        scene->children.push_back(animate_moving(translate(WorldBuilder::new_obj_sphere(0.5f, mat_red), Dir3(-2.0f, 1.5f, 0.0f)), Dir3(2.0f, 0.0f, 0.0f)));
        scene->children.push_back(animate_moving(translate(WorldBuilder::new_obj_sphere(0.5f, mat_blue), Dir3(0.0f, 0.5f, 0.0f)), Dir3(0.0f, 1.0f, 0.0f)));
        finish(*bw, scene, bg_sky(), cb.build());
        bw->default_width = 380; bw->default_height = 320;
    } else if (name == "demo:simple_plane") {  // demo_worlds.rs:298-324
        Camera cam = cam_lookat(60.0f, 9.0f / 16.0f, Point3(0.0f, 6.0f, 10.0f), Point3(0, 0, 0));
        const Material* mat_emit = wb.material_diffuse_light_solid(100.0f * Color(1.0f, 1.0f, 1.0f));
        const Material* mat_floor = wb.material_lambert_solid(Color(0.0f, 0.0f, 0.4f));
        scene->children.push_back(set_poi(WorldBuilder::new_obj_rect(RECT_XY, Point3(0.0f, 2.0f, 0.0f), 1.0f, 1.0f, mat_emit)));
        scene->children.push_back(WorldBuilder::new_obj_rect(RECT_XZ, Point3(0, 0, 0), 10.0f, 10.0f, mat_floor));
        finish(*bw, scene, bg_solid(Color(0.1f, 0.1f, 0.1f)), cam);
        bw->default_width = 640; bw->default_height = 360;
    } else if (name == "demo:earth_mapped") {  // demo_worlds.rs:326-349
        Camera cam = cam_lookat(60.0f, 16.0f / 19.0f, Point3(0.0f, 2.0f, 10.0f), Point3(0, 0, 0));
        const Material* mat_earth = wb.material_lambert(wb.texture_image(assets.earth_w, assets.earth_h, assets.earth_rgb));
        scene->children.push_back(WorldBuilder::new_obj_sphere(2.0f, mat_earth));
        finish(*bw, scene, bg_sky(), cam);
        bw->default_width = 380; bw->default_height = 320;
    } else if (name == "demo:defocus_blur") {  // demo_worlds.rs:465-508
        Dir3 BACKWARD(0, 0, 1), UP(0, 1, 0), RIGHT(1, 0, 0), LEFT(-1, 0, 0), FORWARD(0, 0, -1), DOWN(0, -1, 0);
        CameraBuilder cb;
        cb.vertical_fov(60.0f, 9.0f / 16.0f)
            .set_position(Point3(0, 0, 0) + BACKWARD * 3.0f + UP * 3.0f + 3.0f * RIGHT)
            .look_at_focus(DIR_UP, Point3(0, 0, 0) + FORWARD)
            .set_aperture(0.1f);
        scene->children.push_back(translate(WorldBuilder::new_obj_sphere(100.0f, wb.material_lambert_solid(Color(0.8f, 0.8f, 0.0f))), DOWN * 100.5f));
        scene->children.push_back(translate(WorldBuilder::new_obj_sphere(0.5f, wb.material_lambert_solid(Color(0.7f, 0.3f, 0.3f))), FORWARD));
        scene->children.push_back(translate(WorldBuilder::new_obj_sphere(0.5f, wb.material_metal_solid(Color(0.6f, 0.6f, 0.8f), 0.05f)), LEFT + FORWARD));
        scene->children.push_back(translate(WorldBuilder::new_obj_sphere(0.5f, wb.material_metal_solid(Color(0.8f, 0.6f, 0.2f), 0.5f)), RIGHT + FORWARD));
        scene->children.push_back(translate(WorldBuilder::new_obj_sphere(0.3f, wb.material_dielectric(1.5f)), 0.5f * LEFT + 0.3f * UP));
        finish(*bw, scene, bg_sky(), cb.build());
        bw->default_width = 640; bw->default_height = 360;
    } else if (name == "config:c2") {
        // two suzanne instances (Lambert / Metal) + ground sphere, 1920x1080
        Camera cam = cam_lookat(40.0f, 1080.0f / 1920.0f, Point3(0.0f, 2.0f, 10.0f), Point3(0, 0, 0));
        std::vector<TriangleGeometry> tris = load_obj_file(assets.suzanne_obj);
        const Material* mat_ground = wb.material_lambert_solid(Color(0.4f, 0.4f, 0.4f));
        const Material* mat_lambert = wb.material_lambert_solid(Color(1.0f, 0.2f, 0.2f));
        const Material* mat_metal = wb.material_metal_solid(Color(0.8f, 0.6f, 0.2f), 0.1f);
        scene->children.push_back(translate(WorldBuilder::new_mesh(tris, mat_lambert), Dir3(-2.5f, 0.0f, 0.0f)));
        scene->children.push_back(translate(WorldBuilder::new_mesh(tris, mat_metal), Dir3(2.5f, 0.0f, 0.0f)));
        scene->children.push_back(translate(WorldBuilder::new_obj_sphere(1000.1f, mat_ground), Dir3(0.0f, -1000.0f, 0.0f)));
        finish(*bw, scene, bg_sky(), cam);
        bw->default_width = 1920; bw->default_height = 1080;
    } else if (name == "config:c3") {
        // Cornell rects + POI light + earth-textured sphere + cube.obj (x80), 1024x1024
        float hsize = 278.0f;
        Camera cam = cam_lookat(40.0f, 1.0f, Point3(hsize, hsize, -800.0f), Point3(hsize, hsize, 0.0f));
        const Material* white = wb.material_lambert_solid(Color(0.73f, 0.73f, 0.73f));
        const Material* mat_earth = wb.material_lambert(wb.texture_image(assets.earth_w, assets.earth_h, assets.earth_rgb));
        scene->children.push_back(cornell_box_node(wb, 130.0f));
        scene->children.push_back(translate(WorldBuilder::new_obj_sphere(100.0f, mat_earth), Dir3(190.0f, 100.0f, 190.0f)));
        scene->children.push_back(translate(rotate_up(WorldBuilder::new_mesh(load_obj_file(assets.cube_obj), white, 80.0f), 18.0f), Dir3(380.0f, 80.0f, 350.0f)));
        finish(*bw, scene, bg_solid(Color(0, 0, 0)), cam);
        bw->default_width = 1024; bw->default_height = 1024;
    } else if (name == "config:c4") {
        // 1033 baked suzanne instances (999 944 triangles) + ground sphere, 3840x2160
        Camera cam = cam_lookat(50.0f, 2160.0f / 3840.0f, Point3(0.0f, 12.0f, 14.0f), Point3(0.0f, 0.0f, -40.0f));
        std::vector<TriangleGeometry> tris = load_obj_file(assets.suzanne_obj);
        scene->children.push_back(translate(WorldBuilder::new_obj_sphere(1000.0f, wb.material_lambert_solid(Color(0.5f, 0.5f, 0.5f))), Dir3(0.0f, -1000.0f, 0.0f)));
        for (int i = 0; i < 1033; ++i) {
            float theta = rng.gen_f32() * 360.0f;
            float m = rng.gen_f32();
            const Material* mat;
            if (m < 0.5f) mat = wb.material_lambert_solid(gen_color(rng));
            else { Color c = gen_color(rng); float fz = rng.gen_range(0.0f, 0.3f); mat = wb.material_metal_solid(c, fz); }
            Dir3 pos((float)(i % 33 - 16) * 3.0f, 0.0f, (float)(i / 33) * -3.0f);
            scene->children.push_back(translate(rotate_up(WorldBuilder::new_mesh(tris, mat), theta), pos));
        }
        finish(*bw, scene, bg_sky(), cam);
        bw->default_width = 3840; bw->default_height = 2160;
    } else if (name == "config:c5") {
        // 81x81 random-sphere field + 64 suzanne instances, 3840x2160
        CameraBuilder cb;
        cb.vertical_fov(60.0f, 2160.0f / 3840.0f).set_position(Point3(13.0f, 2.0f, 3.0f)).look_at(DIR_UP, Point3(0, 0, 0))
            .set_focus_distance(10.0f).set_aperture(0.1f);
        add_random_sphere_field(wb, rng, scene, 40);
        std::vector<TriangleGeometry> tris = load_obj_file(assets.suzanne_obj);
        for (int i = 0; i < 64; ++i) {
            float theta = rng.gen_f32() * 360.0f;
            float m = rng.gen_f32();
            const Material* mat;
            if (m < 0.5f) mat = wb.material_lambert_solid(gen_muted_color(rng));
            else { Color c = gen_color(rng); float fz = rng.gen_range(0.0f, 0.3f); mat = wb.material_metal_solid(c, fz); }
            Dir3 pos(((float)(i % 8) - 3.5f) * 9.0f, 0.0f, ((float)(i / 8) - 3.5f) * 9.0f);
            scene->children.push_back(translate(rotate_up(WorldBuilder::new_mesh(tris, mat), theta), pos));
        }
        finish(*bw, scene, bg_sky(), cb.build());
        bw->default_width = 3840; bw->default_height = 2160;
    } else {
        throw std::runtime_error("oracle: unknown world " + name);
    }
    return bw;
}

}  // namespace orc
