// rtiaw_app.hpp — the callers and data formats either side of the hot path, host side:
//   * scene RNG            : rand_xoshiro Xoroshiro128PlusPlus + the rand 0.8.5 draws the demo worlds use (src/app/main.rs:24)
//   * OBJ ingest           : src/app/obj_loader.rs, bug-compatible fan triangulation (Q4)
//   * WorldBuilder / NodeBuilder / NodeRef::finish : src/app/worlds/world_builder.rs
//   * worlds               : the in-scope demo worlds of src/app/worlds/demo_worlds.rs + the BASELINE.json configs C1..C5
// Rows f2 and f3 are in: boxes, run-time Transformation / Animation wrappers (Q2 / Q3 kept), motion blur, Checker and
// Marble (Perlin) textures, constant-density volumes with the Isotropic phase function — all ten demo worlds build.
#pragma once
#include <fstream>
#include <functional>
#include <numeric>
#include <sstream>

#include "rtiaw.hpp"

namespace rtiaw {

// ------------------------------------------------------------------ scene RNG
class SceneRng {  // TRng = Xoroshiro128PlusPlus (src/lib/common.rs:1), seeded like src/app/main.rs:24
public:
    SceneRng() {
        uint8_t seed[16];
        for (int i = 0; i < 16; ++i) seed[i] = (uint8_t)(i + 1);
        std::memcpy(&a_, seed, 8);
        std::memcpy(&b_, seed + 8, 8);
    }
    uint64_t next_u64() {
        const uint64_t s0 = a_, s1 = b_ ^ a_;
        const uint64_t out = rotl(a_ + b_, 17) + a_;
        a_ = rotl(s0, 49) ^ s1 ^ (s1 << 21);
        b_ = rotl(s1, 28);
        return out;
    }
    uint32_t next_u32() { return (uint32_t)next_u64(); }
    float gen_f32() { return (float)(next_u32() >> 8) * (1.0f / 16777216.0f); }  // rand Standard
    Color gen_color() { float r = gen_f32(), g = gen_f32(), b = gen_f32(); return {r, g, b}; }  // gen::<[f32; 3]>()
    float gen_range(float low, float high) {  // rand UniformFloat::sample_single
        float scale = high - low;
        for (;;) {
            uint32_t bits = (next_u32() >> 9) | 0x3F800000u;
            float v12;
            std::memcpy(&v12, &bits, 4);
            const float res = v12 * scale + (low - scale);
            if (res < high) return res;
            uint32_t sb;
            std::memcpy(&sb, &scale, 4);
            --sb;
            std::memcpy(&scale, &sb, 4);
        }
    }

    // rand Uniform::new(-1., 1.).sample (UniformFloat, scale nudged below the open upper bound) — used by rand_distr
    float uniform_pm1() {
        float scale = 2.0f;
        const float max_rand = 1.0f - 1.1920929e-7f;
        while (scale * max_rand + -1.0f >= 1.0f) { uint32_t sb; std::memcpy(&sb, &scale, 4); --sb; std::memcpy(&scale, &sb, 4); }
        uint32_t bits = (next_u32() >> 9) | 0x3F800000u;
        float v12;
        std::memcpy(&v12, &bits, 4);
        return (v12 - 1.0f) * scale + -1.0f;
    }
    Dir3 unit_sphere() {  // rand_distr::UnitSphere (Marsaglia)
        for (;;) {
            const float x1 = uniform_pm1(), x2 = uniform_pm1();
            const float sum = x1 * x1 + x2 * x2;
            if (sum >= 1.0f) continue;
            const float factor = 2.0f * std::sqrt(1.0f - sum);
            return {x1 * factor, x2 * factor, 1.0f - 2.0f * sum};
        }
    }
    uint32_t gen_index(uint32_t ubound) {  // rand UniformInt<u32>::sample_single(0, ubound)
        if (ubound == 0) return next_u32();
        const uint32_t zone = (ubound << __builtin_clz(ubound)) - 1u;
        for (;;) {
            const uint64_t m = (uint64_t)next_u32() * ubound;
            if ((uint32_t)m <= zone) return (uint32_t)(m >> 32);
        }
    }
    template <class T>
    void shuffle(std::vector<T>& v) {  // rand::seq::SliceRandom::shuffle
        for (size_t i = v.size() - 1; i >= 1; --i) std::swap(v[i], v[gen_index((uint32_t)(i + 1))]);
    }
    PerlinData perlin(size_t bits) {  // Perlin::new (perlin.rs:19-36)
        PerlinData p;
        const size_t n = (size_t)1 << bits;
        for (size_t i = 0; i < n; ++i) { const Dir3 g = unit_sphere(); p.ranvec.push_back(g.x); p.ranvec.push_back(g.y); p.ranvec.push_back(g.z); }
        for (std::vector<uint32_t>* perm : {&p.perm_x, &p.perm_y, &p.perm_z}) {
            perm->resize(n);
            std::iota(perm->begin(), perm->end(), 0u);
            shuffle(*perm);
        }
        return p;
    }

private:
    static uint64_t rotl(uint64_t v, int k) { return (v << k) | (v >> (64 - k)); }
    uint64_t a_, b_;
};

// ------------------------------------------------------------------ OBJ ingest (src/app/obj_loader.rs)
class ObjLine {  // the reference's char scanner, over one line
public:
    explicit ObjLine(const std::string& s) : s_(s) {}
    bool done() const { return i_ >= s_.size(); }
    bool accept(char c) {
        if (!done() && s_[i_] == c) { ++i_; return true; }
        return false;
    }
    void blanks() { while (accept(' ')) {} }
    void blanks1() {
        if (!accept(' ')) throw std::runtime_error("obj: ExpectedAtLeastOnWhitespace");
        blanks();
    }
    size_t index() {  // take_usize() - 1
        size_t b = i_;
        while (!done() && std::isdigit((unsigned char)s_[i_])) ++i_;
        if (b == i_) throw std::runtime_error("obj: ExpectedDigits");
        return (size_t)std::stoull(s_.substr(b, i_ - b)) - 1;
    }
    float number() {  // '-'? digits ('.' digits)?
        size_t b = i_;
        accept('-');
        size_t d = i_;
        while (!done() && std::isdigit((unsigned char)s_[i_])) ++i_;
        if (d == i_) throw std::runtime_error("obj: ExpectedDigits");
        if (accept('.')) {
            d = i_;
            while (!done() && std::isdigit((unsigned char)s_[i_])) ++i_;
            if (d == i_) throw std::runtime_error("obj: ExpectedDigits");
        }
        return std::strtof(s_.substr(b, i_ - b).c_str(), nullptr);  // correctly rounded, like str::parse::<f32>
    }

private:
    const std::string& s_;
    size_t i_ = 0;
};

inline std::vector<TriangleGeometry> load_obj_mesh(std::istream& in) {
    std::vector<Point3> pos;
    std::vector<Dir3> nor;
    std::vector<Vec2f> tex;
    std::vector<TriangleGeometry> out;
    std::string line;
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        ObjLine L(line);
        if (L.accept('v')) {
            if (L.accept('n')) {
                Dir3 v; L.blanks1(); v.x = L.number(); L.blanks1(); v.y = L.number(); L.blanks1(); v.z = L.number();
                nor.push_back(v);
            } else if (L.accept('t')) {
                Vec2f v; L.blanks1(); v.x = L.number(); L.blanks1(); v.y = L.number();
                tex.push_back(v);
            } else {
                Point3 v; L.blanks1(); v.x = L.number(); L.blanks1(); v.y = L.number(); L.blanks1(); v.z = L.number();
                pos.push_back(v);
            }
        } else if (L.accept('f')) {
            L.blanks1();
            TriangleGeometry tri{};  // ORIGIN positions, ZERO normals / uvs (obj_loader.rs:227-231)
            for (size_t corner = 0; !L.done(); ++corner) {
                const size_t ip = L.index();
                long it = -1, in_ = -1;
                L.blanks();
                if (L.accept('/')) {
                    L.blanks();
                    if (L.accept('/')) { L.blanks(); in_ = (long)L.index(); }
                    else {
                        it = (long)L.index();
                        L.blanks();
                        if (L.accept('/')) { L.blanks(); in_ = (long)L.index(); }
                    }
                }
                L.blanks();
                if (ip >= pos.size()) throw std::runtime_error("obj: PosIdOutOfRange");
                if (it >= 0 && (size_t)it >= tex.size()) throw std::runtime_error("obj: TexIdOutOfRange");
                if (in_ < 0) throw std::runtime_error("obj: face vertex without a normal (the reference hits todo!())");
                if ((size_t)in_ >= nor.size()) throw std::runtime_error("obj: NorIdOutOfRange");
                const Vec2f uv = it >= 0 ? tex[(size_t)it] : Vec2f{};
                if (corner < 2) {
                    tri.positions[corner] = pos[ip]; tri.normals[corner] = nor[(size_t)in_]; tri.texture_coords[corner] = uv;
                } else {
                    // Q4 (obj_loader.rs:251-264): slot 2 is copied into slot 1 BEFORE the new corner lands in slot 2, so the
                    // first triangle of every face is (v0, ORIGIN, v2) with a zero middle normal.  Kept on purpose.
                    tri.positions[1] = tri.positions[2]; tri.normals[1] = tri.normals[2]; tri.texture_coords[1] = tri.texture_coords[2];
                    tri.positions[2] = pos[ip]; tri.normals[2] = nor[(size_t)in_]; tri.texture_coords[2] = uv;
                    out.push_back(tri);
                }
            }
        }
    }
    return out;
}
inline std::vector<TriangleGeometry> load_obj_file(const std::string& path) {
    std::ifstream f(path);
    if (!f) throw std::runtime_error("cannot open " + path);
    return load_obj_mesh(f);
}

// ------------------------------------------------------------------ node tree (world_builder.rs:194-329)
struct Node;
using NodeRef = std::shared_ptr<Node>;
struct Node {
    struct Geo {
        Geometry geometry;
        MaterialId material;
        bool is_poi;
        float density = 1.0f;  // < 1 -> VolumeGeometry (world_builder.rs:305-309)
    };
    std::vector<Geo> geo;
    Transformation transformation;
    Dir3 moving_animation;
    std::vector<NodeRef> children;
};

class NodeBuilder {
public:
    explicit NodeBuilder(NodeRef n) : n_(std::move(n)) {}
    NodeBuilder& add(const NodeRef& child) { n_->children.push_back(child); return *this; }
    NodeBuilder& add(NodeBuilder& child) { n_->children.push_back(child.build()); return *this; }
    NodeBuilder& add(NodeBuilder&& child) { n_->children.push_back(child.build()); return *this; }
    NodeBuilder& set_all_geo_as_poi() { for (auto& g : n_->geo) g.is_poi = true; return *this; }
    NodeBuilder& rotate_around_up(float deg) { n_->transformation = n_->transformation.rotate_around_up(deg); return *this; }
    NodeBuilder& translate(Dir3 o) { n_->transformation = n_->transformation.translate(o); return *this; }
    NodeBuilder& set_all_geo_densitity(float density) {  // world_builder.rs:254-263 (spelling as in the reference)
        if (!(density >= 0.0f && density < 1.0f)) throw std::invalid_argument("Invalid densitity");
        for (auto& g : n_->geo) g.density = density;
        return *this;
    }
    NodeBuilder& animate_moving(Dir3 velocity) { n_->moving_animation = n_->moving_animation + velocity; return *this; }
    NodeRef build() { return n_; }

private:
    NodeRef n_;
};

class WorldBuilder {
public:
    std::shared_ptr<MaterialTable> tables = std::make_shared<MaterialTable>();
    TextureId texture_solid(Color c) { return tables->texture_solid(c); }
    TextureId texture_image(uint32_t w, uint32_t h, const uint8_t* rgb8) { return tables->texture_image(w, h, rgb8); }
    TextureId texture_checker(float inv_frequency, TextureId even, TextureId odd) { return tables->texture_checker(inv_frequency, even, odd); }
    MaterialId material_lambert_solid(Color c) { return tables->lambert(texture_solid(c)); }
    MaterialId material_lambert(TextureId t) { return tables->lambert(t); }
    MaterialId material_metal_solid(Color c, float fuzz) { return tables->metal(texture_solid(c), fuzz); }
    MaterialId material_dielectric(float ior) { return tables->dielectric(ior); }
    MaterialId material_isotropic_solid(Color c) { return tables->isotropic(texture_solid(c)); }
    TextureId texture_marble(float scale, SceneRng& rng) { return tables->texture_marble(scale, rng.perlin(8)); }  // world_builder.rs:32-37
    MaterialId material_diffuse_light_solid(Color c) { return tables->diffuse_light(texture_solid(c)); }

    static Geometry geo_rect(RectPlane plane, Point3 center, float s0, float s1) {  // world_builder.rs:77-86
        int a0, a1, n;
        rect_axes(plane, a0, a1, n);
        RectGeometry r;
        r.rect_plane = plane;
        r.dist = center[n];
        r.r0[0] = center[a0] - s0 * 0.5f; r.r0[1] = center[a0] + s0 * 0.5f;
        r.r1[0] = center[a1] - s1 * 0.5f; r.r1[1] = center[a1] + s1 * 0.5f;
        return Geometry::Rect(r);
    }
    NodeBuilder new_group() const { return NodeBuilder(std::make_shared<Node>()); }
    NodeBuilder new_obj(const Geometry& g, MaterialId m) const {
        auto n = std::make_shared<Node>();
        n->geo.push_back({g, m, false, 1.0f});
        return NodeBuilder(n);
    }
    NodeBuilder new_obj_box(float width, float height, float depth, MaterialId m) const {  // geo_box, world_builder.rs:88-93
        return new_obj(Geometry::AxisAlignedBox(Aabb{Point3{}, Point3{width, height, depth}}), m);
    }
    NodeBuilder new_obj_sphere(float radius, MaterialId m) const { return new_obj(Geometry::Sphere({Point3{}, radius}), m); }
    NodeBuilder new_obj_sphere_ground(float radius, float height, MaterialId m) const {
        NodeBuilder b = new_obj_sphere(radius, m);
        b.translate({0.0f, height - radius, 0.0f});
        return b;
    }
    NodeBuilder new_obj_rect_xy(Point3 c, float s0, float s1, MaterialId m) const { return new_obj(geo_rect(RectPlane::Xy, c, s0, s1), m); }
    NodeBuilder new_obj_rect_xz(Point3 c, float s0, float s1, MaterialId m) const { return new_obj(geo_rect(RectPlane::Xz, c, s0, s1), m); }
    NodeBuilder new_obj_rect_yz(Point3 c, float s0, float s1, MaterialId m) const { return new_obj(geo_rect(RectPlane::Yz, c, s0, s1), m); }
    // new_mesh_from_file_obj_uniform_material (world_builder.rs:172-191).  `scale` is an ingest-time extension
    // (Transformation has no scale, transformations.rs:4-8): positions are multiplied before any transform.
    NodeBuilder new_mesh_uniform_material(const std::vector<TriangleGeometry>& tris, MaterialId m, float scale = 1.0f) const {
        auto n = std::make_shared<Node>();
        n->geo.reserve(tris.size());
        for (TriangleGeometry t : tris) {
            if (scale != 1.0f) for (auto& p : t.positions) p = p * scale;
            n->geo.push_back({Geometry::Triangle(t), m, false, 1.0f});
        }
        return NodeBuilder(n);
    }
    NodeBuilder new_mesh_from_file_obj_uniform_material(const std::string& path, MaterialId m) const {
        return new_mesh_uniform_material(load_obj_file(path), m);
    }

    // NodeRef::finish (world_builder.rs:273-290): depth-first flatten, bake transforms, first POI rect becomes the
    // light-sampling provider, ONE top-level BVH over all elements.
    std::unique_ptr<World> finish(const NodeRef& root, BackgroundColor background, const Camera& camera) const {
        std::vector<SurfaceGeometry> all;
        auto world = std::make_unique<World>();
        std::function<void(const Node&, const Transformation&)> walk = [&](const Node& node, const Transformation& parent) {
            const Transformation full = parent.then(node.transformation);
            for (const Node::Geo& g : node.geo) {
                SurfaceGeometry e;
                e.geometry = g.geometry.partial_apply_transformation(full, &e.has_transformation, &e.transformation);
                e.material = g.material;
                e.has_animation = !(node.moving_animation == dir::ZERO);
                e.velocity = node.moving_animation;
                e.is_volume = g.density < 1.0f;
                if (e.is_volume) e.neg_inv_density = -1.0f / g.density;
                const Geometry& baked = e.geometry;
                all.push_back(e);
                if (g.is_poi && !e.has_transformation && baked.kind == RT_GEO_RECT && !world->has_scattering_distribution_provider) {
                    world->has_scattering_distribution_provider = true;
                    world->scattering_distribution_provider.rect = baked.rect;
                }
            }
            for (const NodeRef& c : node.children) walk(*c, full);
        };
        walk(*root, Transformation());
        world->camera = camera;
        world->background = background;
        world->tables = tables;
        world->hittable = std::make_shared<Scene>(std::move(all));
        return world;
    }
};

// ------------------------------------------------------------------ worlds
struct Assets {
    std::string suzanne_obj, cube_obj;
    uint32_t earth_w = 0, earth_h = 0;
    const uint8_t* earth_rgb8 = nullptr;  // decoded earthmap.jpg, row 0 = top
};
struct NamedWorld {
    std::unique_ptr<World> world;
    Size2i default_size;
};

namespace detail {
// create_world_final_scene1's sphere field (demo_worlds.rs:407-460); `half` = 11 in the reference.
inline void sphere_field(WorldBuilder& wb, SceneRng& rng, NodeBuilder& scene, int half) {
    const MaterialId mat_ground = wb.material_lambert_solid({0.5f, 0.5f, 0.5f});
    const float ground_radius = 1000.0f;
    const Dir3 ground_center{0.0f, -ground_radius, 0.0f};
    scene.add(wb.new_obj_sphere_ground(ground_radius, 0.0f, mat_ground));
    const MaterialId glass = wb.material_dielectric(1.5f);
    for (int a = -half; a <= half; ++a) {
        for (int b = -half; b <= half; ++b) {
            const float jx = rng.gen_f32() * 0.9f;
            const float jz = rng.gen_f32() * 0.9f;
            const Dir3 center = Dir3{(float)a, 0.2f, (float)b} + Dir3{jx, 0.0f, jz};
            if (!(length(center - Dir3{4.0f, 0.2f, 0.0f}) > 0.9f)) continue;
            const float pick = rng.gen_f32();
            MaterialId mat;
            if (pick < 0.8f) {
                const Color c1 = rng.gen_color();
                const Color c2 = rng.gen_color();
                mat = wb.material_lambert_solid(convolution(c1, c2));
            } else if (pick < 0.95f) {
                const Color c = rng.gen_color();
                const float fuzz = rng.gen_range(0.0f, 0.5f);
                mat = wb.material_metal_solid(c, fuzz);
            } else {
                mat = glass;
            }
            const float small_radius = 0.2f;
            const Dir3 real_center = ground_center + with_length(center - ground_center, ground_radius + small_radius);
            scene.add(wb.new_obj_sphere(small_radius, mat).translate(real_center));
        }
    }
    scene.add(wb.new_obj_sphere(1.0f, glass).translate({0.0f, 1.0f, 0.0f}));
    scene.add(wb.new_obj_sphere(1.0f, wb.material_lambert_solid({0.4f, 0.2f, 0.1f})).translate({-4.0f, 1.0f, 0.0f}));
    scene.add(wb.new_obj_sphere(1.0f, wb.material_metal_solid({0.7f, 0.6f, 0.5f}, 0.0f)).translate({4.0f, 1.0f, 0.0f}));
}
// create_cornell_box_node (demo_worlds.rs:216-264)
inline NodeRef cornell_box_node(WorldBuilder& wb, float light_size) {
    const MaterialId red = wb.material_lambert_solid({0.65f, 0.05f, 0.05f});
    const MaterialId white = wb.material_lambert_solid({0.73f, 0.73f, 0.73f});
    const MaterialId green = wb.material_lambert_solid({0.12f, 0.45f, 0.15f});
    const MaterialId light = wb.material_diffuse_light_solid({15.0f, 15.0f, 15.0f});
    const float h = 278.0f;
    NodeBuilder g = wb.new_group();
    g.add(wb.new_obj_rect_yz({0.0f, h, h}, 2.0f * h, 2.0f * h, red));
    g.add(wb.new_obj_rect_yz({2.0f * h, h, h}, 2.0f * h, 2.0f * h, green));
    g.add(wb.new_obj_rect_xz({h, 0.0f, h}, 2.0f * h, 2.0f * h, white));
    g.add(wb.new_obj_rect_xz({h, 2.0f * h, h}, 2.0f * h, 2.0f * h, white));
    g.add(wb.new_obj_rect_xy({h, h, 2.0f * h}, 2.0f * h, 2.0f * h, white));
    g.add(wb.new_obj_rect_xz({h, 2.0f * h - 1.0f, h}, light_size, light_size, light).set_all_geo_as_poi());
    return g.build();
}
inline MaterialId random_instance_material(WorldBuilder& wb, SceneRng& rng, bool muted) {
    const float pick = rng.gen_f32();
    if (pick < 0.5f) {
        if (muted) {
            const Color c1 = rng.gen_color();
            const Color c2 = rng.gen_color();
            return wb.material_lambert_solid(convolution(c1, c2));
        }
        return wb.material_lambert_solid(rng.gen_color());
    }
    const Color c = rng.gen_color();
    const float fuzz = rng.gen_range(0.0f, 0.3f);
    return wb.material_metal_solid(c, fuzz);
}
}  // namespace detail

// World names: "demo:suzanne", "demo:final_scene1", "demo:simple_plane", "demo:earth_mapped", "demo:defocus_blur"
// (src/app/main.rs:25-37, in-scope subset) and "config:c1" .. "config:c5" (BASELINE.json; DESIGN.md "Workload definitions").
inline NamedWorld create_world(const std::string& name, const Assets& assets) {
    WorldBuilder wb;
    SceneRng rng;
    NodeBuilder scene = wb.new_group();
    NamedWorld out;
    auto sky = BackgroundColor::Sky();
    auto earth_material = [&]() {
        if (!assets.earth_rgb8) throw std::runtime_error("this world needs the decoded earthmap image");
        return wb.material_lambert(wb.texture_image(assets.earth_w, assets.earth_h, assets.earth_rgb8));
    };
    if (name == "demo:suzanne") {  // demo_worlds.rs:6-28
        const Camera cam = Camera::build().vertical_fov(40.0f, 3.0f / 4.0f).position({0.0f, 2.0f, 10.0f}).look_at(dir::UP, {0, 0, 0}).build();
        const MaterialId mat_ground = wb.material_lambert_solid({0.4f, 0.4f, 0.4f});
        const MaterialId mat_monkey = wb.material_lambert_solid({1.0f, 0.2f, 0.2f});
        scene.add(wb.new_mesh_from_file_obj_uniform_material(assets.suzanne_obj, mat_monkey));
        scene.add(wb.new_obj_sphere(1000.1f, mat_ground).translate({0.0f, -1000.0f, 0.0f}));
        out.world = wb.finish(scene.build(), sky, cam);
        out.default_size = {400, 300};
    } else if (name == "demo:final_scene1" || name == "config:c1") {  // demo_worlds.rs:395-463
        const bool c1 = name == "config:c1";
        const Camera cam = Camera::build().vertical_fov(60.0f, c1 ? 800.0f / 1200.0f : 9.0f / 16.0f).position({13.0f, 2.0f, 3.0f})
                               .look_at(dir::UP, {0, 0, 0}).focus_distance(10.0f).aperture(0.1f).build();
        detail::sphere_field(wb, rng, scene, 11);
        out.world = wb.finish(scene.build(), sky, cam);
        out.default_size = {1200, c1 ? 800 : 675};
    } else if (name == "demo:cornell_box") {  // demo_worlds.rs:266-296
        const float h = 278.0f;
        const Camera cam = Camera::build().vertical_fov(40.0f, 1.0f).position({h, h, -800.0f}).look_at(dir::UP, {h, h, 0.0f}).build();
        const MaterialId white = wb.material_lambert_solid({0.73f, 0.73f, 0.73f});
        scene.add(detail::cornell_box_node(wb, 130.0f));
        scene.add(wb.new_obj_box(165.0f, 330.0f, 165.0f, white).rotate_around_up(15.0f).translate({265.0f, 0.0f, 295.0f}));
        scene.add(wb.new_obj_box(165.0f, 165.0f, 165.0f, white).rotate_around_up(-18.0f).translate({130.0f, 0.0f, 65.0f}));
        out.world = wb.finish(scene.build(), BackgroundColor::Solid({0, 0, 0}), cam);
        out.default_size = {600, 600};
    } else if (name == "demo:cornell_box_smoke") {  // demo_worlds.rs:179-214
        const float h = 278.0f;
        const Camera cam = Camera::build().vertical_fov(40.0f, 1.0f).position({h, h, -800.0f}).look_at(dir::UP, {h, h, 0.0f}).build();
        const MaterialId smoke_black = wb.material_isotropic_solid({0.0f, 0.0f, 0.0f});
        const MaterialId smoke_white = wb.material_isotropic_solid({1.0f, 1.0f, 1.0f});
        scene.add(detail::cornell_box_node(wb, 130.0f));
        scene.add(wb.new_obj_box(165.0f, 330.0f, 165.0f, smoke_white).rotate_around_up(15.0f).translate({265.0f, 0.0f, 295.0f}).set_all_geo_densitity(0.01f));
        scene.add(wb.new_obj_box(165.0f, 165.0f, 165.0f, smoke_black).rotate_around_up(-18.0f).translate({130.0f, 0.0f, 65.0f}).set_all_geo_densitity(0.01f));
        out.world = wb.finish(scene.build(), BackgroundColor::Solid({0, 0, 0}), cam);
        out.default_size = {600, 600};
    } else if (name == "demo:perlin_spheres") {  // demo_worlds.rs:154-177
        const Camera cam = Camera::build().vertical_fov(40.0f, 3.0f / 4.0f).position({13.0f, 2.0f, 3.0f}).look_at(dir::UP, {0, 0, 0}).build();
        const MaterialId mat_noise = wb.material_lambert(wb.texture_marble(4.0f, rng));
        scene.add(wb.new_obj_sphere_ground(1000.0f, 0.0f, mat_noise));
        scene.add(wb.new_obj_sphere(2.0f, mat_noise).translate({0.0f, 2.0f, 0.0f}));
        out.world = wb.finish(scene.build(), sky, cam);
        out.default_size = {400, 300};
    } else if (name == "demo:final_scene2") {  // demo_worlds.rs:30-152
        const Camera cam = Camera::build().vertical_fov(40.0f, 1.0f).position({478.0f, 278.0f, -600.0f}).look_at(dir::UP, {278.0f, 278.0f, 0.0f})
                               .motion_blur(0.0f, 1.0f).build();
        const MaterialId mat_ground = wb.material_lambert_solid({0.48f, 0.83f, 0.53f});
        NodeBuilder ground = wb.new_group();
        {
            const int boxes_per_side = 20;
            const Point3 rmin{-1000.0f, 0.0f, -1000.0f}, rmax{1000.0f, 101.0f, 1000.0f};
            const float width = (rmax.x - rmin.x) / (float)boxes_per_side;
            const float height = rmax.y - rmin.y;
            const float depth = (rmax.z - rmin.z) / (float)boxes_per_side;
            for (int i = 0; i < boxes_per_side; ++i)
                for (int j = 0; j < boxes_per_side; ++j) {
                    const Dir3 mn = rmin + Dir3{(float)i * width, 0.0f, (float)j * depth} - Point3{};
                    const float hbox = height * rng.gen_f32();
                    ground.add(wb.new_obj_box(width, hbox, depth, mat_ground).translate(mn));
                }
        }
        const MaterialId mat_light = wb.material_diffuse_light_solid({7.0f, 7.0f, 7.0f});
        NodeBuilder light = wb.new_obj_rect_xz({273.0f, 554.0f, 279.0f}, 300.0f, 300.0f, mat_light);
        const MaterialId mat_glass = wb.material_dielectric(1.5f);
        const MaterialId mat_metal = wb.material_metal_solid({0.8f, 0.8f, 0.9f}, 1.0f);
        const MaterialId mat_marble = wb.material_lambert(wb.texture_marble(0.1f, rng));
        const MaterialId mat_earth = earth_material();
        const MaterialId mat_fog = wb.material_isotropic_solid({0.2f, 0.4f, 0.9f});
        const MaterialId mat_moving = wb.material_lambert_solid({0.7f, 0.3f, 0.1f});
        NodeBuilder floating = wb.new_group();
        floating.add(wb.new_obj_sphere(70.0f, mat_glass).translate({360.0f, 150.0f, 145.0f}));
        floating.add(wb.new_obj_sphere(70.0f, mat_fog).translate({360.0f, 150.0f, 145.0f}).set_all_geo_densitity(0.2f));
        floating.add(wb.new_obj_sphere(50.0f, mat_glass).translate({260.0f, 150.0f, 45.0f}));
        floating.add(wb.new_obj_sphere(50.0f, mat_metal).translate({0.0f, 150.0f, 145.0f}));
        floating.add(wb.new_obj_sphere(80.0f, mat_marble).translate({220.0f, 280.0f, 300.0f}));
        floating.add(wb.new_obj_sphere(100.0f, mat_earth).translate({400.0f, 200.0f, 400.0f}));
        floating.add(wb.new_obj_sphere(50.0f, mat_moving).translate({400.0f, 400.0f, 200.0f}).animate_moving({30.0f, 0.0f, 0.0f}));
        NodeBuilder sphere_cube = wb.new_group();
        {
            const MaterialId material = wb.material_lambert_solid({0.75f, 0.75f, 0.75f});
            const Transformation rot = Transformation().rotate_around_up(15.0f);
            const Dir3 dir_1 = rot.apply_direction(dir::RIGHT), dir_2 = rot.apply_direction(dir::UP), dir_3 = rot.apply_direction(dir::FORWARD);
            const Dir3 origin{-100.0f, 270.0f, 395.0f};
            for (int i = 0; i < 1000; ++i) {
                const float r0 = rng.gen_f32(), r1 = rng.gen_f32(), r2 = rng.gen_f32();
                const Dir3 p = origin + 165.0f * (dir_1 * r0 + dir_2 * r1 + dir_3 * r2);
                sphere_cube.add(wb.new_obj_sphere(10.0f, material).translate(p));
            }
        }
        NodeBuilder obj_fog = wb.new_obj_sphere(5000.0f, wb.material_isotropic_solid({1.0f, 1.0f, 1.0f}));
        obj_fog.set_all_geo_densitity(0.0001f);
        scene.add(obj_fog).add(ground).add(sphere_cube).add(floating).add(light);
        out.world = wb.finish(scene.build(), BackgroundColor::Solid({0, 0, 0}), cam);
        out.default_size = {600, 600};
    } else if (name == "demo:moving_spheres") {  // demo_worlds.rs:351-393
        const Camera cam = Camera::build().vertical_fov(60.0f, 16.0f / 19.0f).position({0.0f, 2.0f, 10.0f}).look_at(dir::UP, {0.0f, 2.0f, 0.0f})
                               .motion_blur(0.0f, 0.5f).build();
        const TextureId tex_black = wb.texture_solid({0.0f, 0.0f, 0.0f});
        const TextureId tex_white = wb.texture_solid({1.0f, 1.0f, 1.0f});
        const MaterialId mat_ground = wb.material_lambert(wb.texture_checker(10.0f, tex_black, tex_white));
        const MaterialId mat_red = wb.material_lambert_solid({0.6f, 0.2f, 0.2f});
        const MaterialId mat_blue = wb.material_lambert_solid({0.2f, 0.2f, 0.6f});
        scene.add(wb.new_obj_sphere(100.0f, mat_ground).translate({0.0f, -100.0f, 0.0f}));
        scene.add(wb.new_obj_sphere(0.5f, mat_red).translate({-2.0f, 1.5f, 0.0f}).animate_moving({2.0f, 0.0f, 0.0f}));
        scene.add(wb.new_obj_sphere(0.5f, mat_blue).translate({0.0f, 0.5f, 0.0f}).animate_moving({0.0f, 1.0f, 0.0f}));
        out.world = wb.finish(scene.build(), sky, cam);
        out.default_size = {380, 320};
    } else if (name == "demo:simple_plane") {  // demo_worlds.rs:298-324
        const Camera cam = Camera::build().vertical_fov(60.0f, 9.0f / 16.0f).position({0.0f, 6.0f, 10.0f}).look_at(dir::UP, {0, 0, 0}).build();
        const MaterialId mat_emit = wb.material_diffuse_light_solid(100.0f * Color{1.0f, 1.0f, 1.0f});
        const MaterialId mat_floor = wb.material_lambert_solid({0.0f, 0.0f, 0.4f});
        scene.add(wb.new_obj_rect_xy({0.0f, 2.0f, 0.0f}, 1.0f, 1.0f, mat_emit).set_all_geo_as_poi());
        scene.add(wb.new_obj_rect_xz({0, 0, 0}, 10.0f, 10.0f, mat_floor));
        out.world = wb.finish(scene.build(), BackgroundColor::Solid({0.1f, 0.1f, 0.1f}), cam);
        out.default_size = {640, 360};
    } else if (name == "demo:earth_mapped") {  // demo_worlds.rs:326-349
        const Camera cam = Camera::build().vertical_fov(60.0f, 16.0f / 19.0f).position({0.0f, 2.0f, 10.0f}).look_at(dir::UP, {0, 0, 0}).build();
        scene.add(wb.new_obj_sphere(2.0f, earth_material()));
        out.world = wb.finish(scene.build(), sky, cam);
        out.default_size = {380, 320};
    } else if (name == "demo:defocus_blur") {  // demo_worlds.rs:465-508
        using namespace dir;
        const Camera cam = Camera::build().vertical_fov(60.0f, 9.0f / 16.0f)
                               .position(Point3{} + BACKWARD * 3.0f + UP * 3.0f + 3.0f * RIGHT)
                               .look_at_focus(UP, Point3{} + FORWARD).aperture(0.1f).build();
        scene.add(wb.new_obj_sphere(100.0f, wb.material_lambert_solid({0.8f, 0.8f, 0.0f})).translate(DOWN * 100.5f));
        scene.add(wb.new_obj_sphere(0.5f, wb.material_lambert_solid({0.7f, 0.3f, 0.3f})).translate(FORWARD));
        scene.add(wb.new_obj_sphere(0.5f, wb.material_metal_solid({0.6f, 0.6f, 0.8f}, 0.05f)).translate(LEFT + FORWARD));
        scene.add(wb.new_obj_sphere(0.5f, wb.material_metal_solid({0.8f, 0.6f, 0.2f}, 0.5f)).translate(RIGHT + FORWARD));
        scene.add(wb.new_obj_sphere(0.3f, wb.material_dielectric(1.5f)).translate(0.5f * LEFT + 0.3f * UP));
        out.world = wb.finish(scene.build(), sky, cam);
        out.default_size = {640, 360};
    } else if (name == "config:c2") {  // two suzanne instances (Lambert / Metal) + ground sphere
        const Camera cam = Camera::build().vertical_fov(40.0f, 1080.0f / 1920.0f).position({0.0f, 2.0f, 10.0f}).look_at(dir::UP, {0, 0, 0}).build();
        const std::vector<TriangleGeometry> tris = load_obj_file(assets.suzanne_obj);
        const MaterialId mat_ground = wb.material_lambert_solid({0.4f, 0.4f, 0.4f});
        const MaterialId mat_lambert = wb.material_lambert_solid({1.0f, 0.2f, 0.2f});
        const MaterialId mat_metal = wb.material_metal_solid({0.8f, 0.6f, 0.2f}, 0.1f);
        scene.add(wb.new_mesh_uniform_material(tris, mat_lambert).translate({-2.5f, 0.0f, 0.0f}));
        scene.add(wb.new_mesh_uniform_material(tris, mat_metal).translate({2.5f, 0.0f, 0.0f}));
        scene.add(wb.new_obj_sphere(1000.1f, mat_ground).translate({0.0f, -1000.0f, 0.0f}));
        out.world = wb.finish(scene.build(), sky, cam);
        out.default_size = {1920, 1080};
    } else if (name == "config:c3") {  // Cornell rects + POI light + earth sphere + cube.obj x80
        const float h = 278.0f;
        const Camera cam = Camera::build().vertical_fov(40.0f, 1.0f).position({h, h, -800.0f}).look_at(dir::UP, {h, h, 0.0f}).build();
        const MaterialId white = wb.material_lambert_solid({0.73f, 0.73f, 0.73f});
        const MaterialId earth = earth_material();
        scene.add(detail::cornell_box_node(wb, 130.0f));
        scene.add(wb.new_obj_sphere(100.0f, earth).translate({190.0f, 100.0f, 190.0f}));
        scene.add(wb.new_mesh_uniform_material(load_obj_file(assets.cube_obj), white, 80.0f).rotate_around_up(18.0f).translate({380.0f, 80.0f, 350.0f}));
        out.world = wb.finish(scene.build(), BackgroundColor::Solid({0, 0, 0}), cam);
        out.default_size = {1024, 1024};
    } else if (name == "config:c4") {  // 1033 baked suzanne instances (999 944 triangles) + ground sphere
        const Camera cam = Camera::build().vertical_fov(50.0f, 2160.0f / 3840.0f).position({0.0f, 12.0f, 14.0f}).look_at(dir::UP, {0.0f, 0.0f, -40.0f}).build();
        const std::vector<TriangleGeometry> tris = load_obj_file(assets.suzanne_obj);
        scene.add(wb.new_obj_sphere(1000.0f, wb.material_lambert_solid({0.5f, 0.5f, 0.5f})).translate({0.0f, -1000.0f, 0.0f}));
        for (int i = 0; i < 1033; ++i) {
            const float theta = rng.gen_f32() * 360.0f;
            const MaterialId mat = detail::random_instance_material(wb, rng, false);
            const Dir3 at{(float)(i % 33 - 16) * 3.0f, 0.0f, (float)(i / 33) * -3.0f};
            scene.add(wb.new_mesh_uniform_material(tris, mat).rotate_around_up(theta).translate(at));
        }
        out.world = wb.finish(scene.build(), sky, cam);
        out.default_size = {3840, 2160};
    } else if (name == "config:c5") {  // 81x81 random-sphere field + 64 suzanne instances
        const Camera cam = Camera::build().vertical_fov(60.0f, 2160.0f / 3840.0f).position({13.0f, 2.0f, 3.0f}).look_at(dir::UP, {0, 0, 0})
                               .focus_distance(10.0f).aperture(0.1f).build();
        detail::sphere_field(wb, rng, scene, 40);
        const std::vector<TriangleGeometry> tris = load_obj_file(assets.suzanne_obj);
        for (int i = 0; i < 64; ++i) {
            const float theta = rng.gen_f32() * 360.0f;
            const MaterialId mat = detail::random_instance_material(wb, rng, true);
            const Dir3 at{((float)(i % 8) - 3.5f) * 9.0f, 0.0f, ((float)(i / 8) - 3.5f) * 9.0f};
            scene.add(wb.new_mesh_uniform_material(tris, mat).rotate_around_up(theta).translate(at));
        }
        out.world = wb.finish(scene.build(), sky, cam);
        out.default_size = {3840, 2160};
    } else {
        throw std::runtime_error("unknown world '" + name + "' (the ten reference demos and config:c1..c5 are available)");
    }
    return out;
}

}  // namespace rtiaw
