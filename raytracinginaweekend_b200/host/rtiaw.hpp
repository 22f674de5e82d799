// rtiaw.hpp — host-side mirror of the reference's `src/lib` API for the render hot path (C++17).
//
// The reference is a Rust crate (`ray_tracing_in_a_weekend`); this image has no Rust toolchain, so the host
// side above the C ABI (include/b200rt.h) is written in C++ with the reference's names and argument meaning:
//   Vec3 / Point3 / Dir3, Vec2f, Color, Ray, Size2i, Aabb, Transformation,
//   SphereGeometry / RectGeometry / TriangleGeometry / Geometry, Texture, Material, SceneElement,
//   BoundingVolumeHierarchy::build (= ::new), Scene, Camera::build() builder, BackgroundColor,
//   WorldScatteringDistributionProvider, World, RenderMode, render().
// Unlike the reference, nothing here intersects or shades on the CPU: Scene::hit and render() flatten the
// scene graph into the SoA arrays of RtSceneDesc and call the sm_100a kernels in libb200rt.so.
// file:line citations refer to the reference repo.
#pragma once
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <numeric>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "b200rt.h"

namespace rtiaw {

// libm calls the reference makes at RUN time (f32::tan / f32::sin_cos -> the platform libm); `volatile` stops the compiler
// from folding them at compile time with a differently-rounded result.
inline float libm_tan(float x) { volatile float v = x; return std::tan(v); }
inline float libm_sin(float x) { volatile float v = x; return std::sin(v); }
inline float libm_cos(float x) { volatile float v = x; return std::cos(v); }

// ------------------------------------------------------------------ math types (src/lib/vec3.rs, vec2.rs, color.rs)
struct Vec3 {
    float x = 0, y = 0, z = 0;
    constexpr Vec3() = default;
    constexpr Vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    float& at(int i) { return i == 0 ? x : (i == 1 ? y : z); }
};
inline Vec3 operator+(Vec3 a, Vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec3 operator-(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec3 operator-(Vec3 a) { return {-a.x, -a.y, -a.z}; }
inline Vec3 operator*(Vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline Vec3 operator*(float s, Vec3 a) { return a * s; }
inline Vec3 operator/(Vec3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline bool operator==(Vec3 a, Vec3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
using Point3 = Vec3;
using Dir3 = Vec3;
using Color = Vec3;
namespace dir {
constexpr Dir3 ZERO{0, 0, 0}, RIGHT{1, 0, 0}, LEFT{-1, 0, 0}, UP{0, 1, 0}, DOWN{0, -1, 0}, FORWARD{0, 0, -1}, BACKWARD{0, 0, 1};
}
inline float dot(Dir3 a, Dir3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }  // vec3.rs:189-191
inline Dir3 cross(Dir3 a, Dir3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline float length(Dir3 a) { return std::sqrt(dot(a, a)); }
inline Dir3 with_length(Dir3 a, float len) { return a * (len / length(a)); }  // vec3.rs:206-208
inline Dir3 unit(Dir3 a) { return with_length(a, 1.0f); }
inline Color convolution(Color a, Color b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }  // color.rs:58-64

struct Vec2f {
    float x = 0, y = 0;
};
struct Size2i {  // src/lib/size2i.rs
    int32_t width = 0, height = 0;
    float aspect_ratio() const { return (float)height / (float)width; }  // :13-15 (height / width!)
    size_t count() const { return (size_t)width * (size_t)height; }
};
struct Ray {  // src/lib/ray.rs
    Point3 origin;
    Dir3 direction;
    float time = 0.0f;
};
struct Range {
    float start, end;
};

// ------------------------------------------------------------------ src/lib/aabb.rs:10-63
struct Aabb {
    Point3 min, max;
    static Aabb surrounding(const Aabb& a, const Aabb& b) {
        return {{std::fmin(a.min.x, b.min.x), std::fmin(a.min.y, b.min.y), std::fmin(a.min.z, b.min.z)},
                {std::fmax(a.max.x, b.max.x), std::fmax(a.max.y, b.max.y), std::fmax(a.max.z, b.max.z)}};
    }
    // Transformation::apply_aabb (hittable.rs:285-291) as it actually behaves (Q2): `for mut c in corners` transforms COPIES, so
    // the result is the bounding box of the eight UNtransformed corners (aabb.rs:178-192) — equal to *this up to one rounding
    // of min + (max - min).
    Aabb apply_aabb_as_reference() const {
        const Dir3 right{max.x - min.x, 0.0f, 0.0f}, up{0.0f, max.y - min.y, 0.0f}, fwd{0.0f, 0.0f, max.z - min.z};
        const Point3 c[8] = {min, min + right, min + up, min + fwd, max, max - right, max - up, max - fwd};
        Aabb b{c[0], c[0]};
        for (int k = 1; k < 8; ++k) b = surrounding(b, Aabb{c[k], c[k]});
        return b;
    }
};

// ------------------------------------------------------------------ src/lib/transformations.rs
class Transformation {
public:
    Dir3 offset;
    float y_sine = 0.0f, y_cosine = 1.0f;
    static void rotate_y(float c, float s, Vec3& v) {  // rotate_around_up :107-111
        const float a = v.x, b = v.z;
        v.x = c * a + s * b;
        v.z = -s * a + c * b;
    }
    bool is_zero() const { return offset == dir::ZERO && y_sine == 0.0f && y_cosine == 1.0f; }
    Transformation translate(Dir3 o) const {
        Transformation r = *this;
        r.offset = r.offset + o;
        return r;
    }
    Transformation rotate_around_up(float angle_deg) const {  // :26-36
        Transformation r = *this;
        const float rad = angle_deg * (3.14159265358979323846f / 180.0f);
        const float s = libm_sin(rad), c = libm_cos(rad);
        rotate_y(c, s, r.offset);
        Vec3 sc(y_sine, 0.0f, y_cosine);
        rotate_y(c, s, sc);
        r.y_sine = sc.x;
        r.y_cosine = sc.z;
        return r;
    }
    Point3 apply_point(Point3 p) const {
        rotate_y(y_cosine, y_sine, p);
        return p + offset;
    }
    Dir3 apply_normal(Dir3 d) const {
        rotate_y(y_cosine, y_sine, d);
        return d;
    }
    Dir3 apply_direction(Dir3 d) const { return apply_normal(d); }
    // split_translation_remainder (:66-70): (offset, the same rotation without the offset)
    Transformation rotation_remainder() const {
        Transformation r = *this;
        r.offset = dir::ZERO;
        return r;
    }
    Transformation then(const Transformation& next) const {  // :89-98
        Transformation r;
        r.offset = next.apply_direction(offset) + next.offset;
        r.y_sine = y_sine * next.y_cosine + y_cosine * next.y_sine;
        r.y_cosine = y_cosine * next.y_cosine - y_sine * next.y_sine;
        return r;
    }
};

// ------------------------------------------------------------------ geometry (src/lib/hittable/*.rs, hittable.rs:104-110)
enum class RectPlane : uint32_t { Xy = RT_RECT_XY, Xz = RT_RECT_XZ, Yz = RT_RECT_YZ };
inline void rect_axes(RectPlane p, int& a0, int& a1, int& n) {  // rect_geometry.rs:14-21
    switch (p) {
        case RectPlane::Xy: a0 = 0; a1 = 1; n = 2; break;
        case RectPlane::Xz: a0 = 0; a1 = 2; n = 1; break;
        default: a0 = 1; a1 = 2; n = 0; break;
    }
}
struct SphereGeometry {
    Point3 center;
    float radius = 0;
};
struct RectGeometry {
    RectPlane rect_plane = RectPlane::Xy;
    float dist = 0;
    float r0[2] = {0, 0}, r1[2] = {0, 0};
};
struct TriangleGeometry {
    Point3 positions[3];
    Dir3 normals[3];
    Vec2f texture_coords[3];
};
struct Geometry {
    uint32_t kind = RT_GEO_SPHERE;
    SphereGeometry sphere;
    RectGeometry rect;
    TriangleGeometry triangle;
    Aabb box;  // Geometry::AxisAlignedBox(Aabb)
    static Geometry AxisAlignedBox(const Aabb& b) { Geometry g; g.kind = RT_GEO_BOX; g.box = b; return g; }
    static Geometry Sphere(SphereGeometry s) {
        if (!(s.radius >= 0.0f)) throw std::invalid_argument("assertion failed: radius >= 0.0 (sphere_geometry.rs:13)");
        Geometry g; g.kind = RT_GEO_SPHERE; g.sphere = s; return g;
    }
    static Geometry Rect(RectGeometry r) { Geometry g; g.kind = RT_GEO_RECT; g.rect = r; return g; }
    static Geometry Triangle(const TriangleGeometry& t) { Geometry g; g.kind = RT_GEO_TRIANGLE; g.triangle = t; return g; }

    Aabb bounding_box() const {  // hittable.rs:150-157
        if (kind == RT_GEO_SPHERE) {
            const Dir3 r(sphere.radius, sphere.radius, sphere.radius);
            return {sphere.center - r, sphere.center + r};
        }
        if (kind == RT_GEO_RECT) {  // rect_geometry.rs:87-98 with thickness 0.01
            int a0, a1, n;
            rect_axes(rect.rect_plane, a0, a1, n);
            Aabb b;
            b.min.at(a0) = rect.r0[0]; b.min.at(a1) = rect.r1[0]; b.min.at(n) = rect.dist - 0.01f;
            b.max.at(a0) = rect.r0[1]; b.max.at(a1) = rect.r1[1]; b.max.at(n) = rect.dist + 0.01f;
            return b;
        }
        if (kind == RT_GEO_BOX) return box;
        Aabb b{triangle.positions[0], triangle.positions[0]};
        for (int k = 1; k < 3; ++k) b = Aabb::surrounding(b, Aabb{triangle.positions[k], triangle.positions[k]});
        return b;
    }
    // hittable.rs:165-208.  Returns the baked geometry; *has_rest / *rest describe the run-time Transformation that remains
    // (rects: the whole transform; boxes: the rotation-only remainder after the translation was baked — Q3).
    Geometry partial_apply_transformation(const Transformation& t, bool* has_rest, Transformation* rest) const {
        Geometry g = *this;
        *has_rest = false;
        if (kind == RT_GEO_SPHERE) {
            g.sphere.center = t.apply_point(sphere.center);
        } else if (kind == RT_GEO_TRIANGLE) {
            for (int k = 0; k < 3; ++k) {
                g.triangle.positions[k] = t.apply_point(triangle.positions[k]);
                g.triangle.normals[k] = t.apply_normal(triangle.normals[k]);
            }
        } else if (kind == RT_GEO_BOX) {
            g.box = Aabb{box.min + t.offset, box.max + t.offset};
            *rest = t.rotation_remainder();
            *has_rest = !rest->is_zero();
        } else {
            *rest = t;
            *has_rest = !t.is_zero();
        }
        return g;
    }
    void write(RtItem& it) const {
        std::memset(it.geom, 0, sizeof(it.geom));
        it.kind = kind;
        float* g = it.geom;
        if (kind == RT_GEO_SPHERE) {
            g[0] = sphere.center.x; g[1] = sphere.center.y; g[2] = sphere.center.z; g[3] = sphere.radius;
        } else if (kind == RT_GEO_RECT) {
            g[0] = (float)(uint32_t)rect.rect_plane; g[1] = rect.dist;
            g[2] = rect.r0[0]; g[3] = rect.r0[1]; g[4] = rect.r1[0]; g[5] = rect.r1[1];
        } else if (kind == RT_GEO_BOX) {
            g[0] = box.min.x; g[1] = box.min.y; g[2] = box.min.z; g[3] = box.max.x; g[4] = box.max.y; g[5] = box.max.z;
        } else {
            for (int k = 0; k < 3; ++k) {
                g[3 * k] = triangle.positions[k].x; g[3 * k + 1] = triangle.positions[k].y; g[3 * k + 2] = triangle.positions[k].z;
                g[9 + 3 * k] = triangle.normals[k].x; g[10 + 3 * k] = triangle.normals[k].y; g[11 + 3 * k] = triangle.normals[k].z;
                g[18 + 2 * k] = triangle.texture_coords[k].x; g[19 + 2 * k] = triangle.texture_coords[k].y;
            }
        }
    }
};

// ------------------------------------------------------------------ textures / materials (tables instead of arena refs)
struct ImageData {
    uint32_t width = 0, height = 0;
    std::vector<uint8_t> rgb8;
};
struct PerlinData {  // Perlin (src/lib/perlin.rs:10-16) with bits = 8
    std::vector<float> ranvec;  // 256 * 3
    std::vector<uint32_t> perm_x, perm_y, perm_z;
};
using TextureId = uint32_t;
using MaterialId = uint32_t;

// Owns the material / texture / image tables a scene refers to (the reference keeps them in a bumpalo arena).
class MaterialTable {
public:
    std::vector<RtTexture> textures;
    std::vector<RtMaterial> materials;
    std::vector<ImageData> images;
    std::vector<PerlinData> perlins;
    TextureId texture_marble(float scale, PerlinData noise) {  // Texture::Marble { scale, noise }
        perlins.push_back(std::move(noise));
        RtTexture t{};
        t.kind = RT_TEX_MARBLE; t.image = (uint32_t)perlins.size() - 1; t.inv_frequency = scale; t.even = t.odd = 0xFFFFFFFFu;
        textures.push_back(t);
        return (TextureId)textures.size() - 1;
    }
    MaterialId isotropic(TextureId albedo) { return add(RT_MAT_ISOTROPIC, albedo, 0.0f); }
    TextureId texture_solid(Color c) {  // Texture::Solid
        RtTexture t{};
        t.kind = RT_TEX_SOLID; t.image = 0xFFFFFFFFu; t.even = t.odd = 0xFFFFFFFFu; t.rgb[0] = c.x; t.rgb[1] = c.y; t.rgb[2] = c.z;
        textures.push_back(t);
        return (TextureId)textures.size() - 1;
    }
    TextureId texture_image(uint32_t w, uint32_t h, const uint8_t* rgb8) {  // Texture::Image
        ImageData im;
        im.width = w; im.height = h; im.rgb8.assign(rgb8, rgb8 + (size_t)w * h * 3);
        images.push_back(std::move(im));
        RtTexture t{};
        t.kind = RT_TEX_IMAGE; t.image = (uint32_t)images.size() - 1; t.even = t.odd = 0xFFFFFFFFu;
        textures.push_back(t);
        return (TextureId)textures.size() - 1;
    }
    TextureId texture_checker(float inv_frequency, TextureId even, TextureId odd) {  // Texture::Checker
        RtTexture t{};
        t.kind = RT_TEX_CHECKER; t.image = 0xFFFFFFFFu; t.inv_frequency = inv_frequency; t.even = even; t.odd = odd;
        textures.push_back(t);
        return (TextureId)textures.size() - 1;
    }
    MaterialId add(uint32_t kind, TextureId tex, float param) {
        materials.push_back(RtMaterial{kind, tex, param});
        return (MaterialId)materials.size() - 1;
    }
    MaterialId lambert(TextureId albedo) { return add(RT_MAT_LAMBERT, albedo, 0.0f); }
    MaterialId metal(TextureId albedo, float fuzz) { return add(RT_MAT_METAL, albedo, fuzz); }
    MaterialId dielectric(float index_of_refraction) { return add(RT_MAT_DIELECTRIC, 0xFFFFFFFFu, index_of_refraction); }
    MaterialId diffuse_light(TextureId emit) { return add(RT_MAT_DIFFUSE_LIGHT, emit, 0.0f); }
};

// ------------------------------------------------------------------ SceneElement / BVH / Scene (hittable.rs:112-138, 345-482)
// SceneElement::SurfaceGeometry(geo, material), optionally wrapped by SceneElement::Transformation and then
// SceneElement::Animation — the nesting NodeRef::finish produces (world_builder.rs:305-315).
struct SurfaceGeometry {
    Geometry geometry;
    MaterialId material = 0;
    bool has_transformation = false;
    Transformation transformation;
    bool has_animation = false;
    Dir3 velocity;
    bool is_volume = false;        // SceneElement::VolumeGeometry: `geometry` is the boundary, `material` the phase function
    float neg_inv_density = 0.0f;  // VolumeGeometry::new (hittable.rs:301-307): -1.0 / density
    // SceneElement::bounding_box (hittable.rs:249-267) incl. Q2: wrappers neither move nor motion-extend the box
    Aabb bounding_box() const {
        Aabb b = geometry.bounding_box();
        if (has_transformation) b = b.apply_aabb_as_reference();
        if (has_animation) b = Aabb::surrounding(b.apply_aabb_as_reference(), b.apply_aabb_as_reference());
        return b;
    }
};

// BoundingVolumeHierarchy::new (hittable.rs:360-427): same topology as the reference — pre-order node numbering,
// stable sort on aabb.min[axis], median split, child axis = (axis + child_len) % 3, leaf id = item index —
// emitted directly as the flat RtBvhNode array the device packer consumes.
class BoundingVolumeHierarchy {
public:
    std::vector<RtBvhNode> nodes;
    uint32_t initial_index = 0;
    uint32_t depth = 0;

    static BoundingVolumeHierarchy build(const std::vector<SurfaceGeometry>& items) {
        if (items.empty()) throw std::invalid_argument("BoundingVolumeHierarchy: no items (the reference recurses forever)");
        BoundingVolumeHierarchy bvh;
        const size_t n = items.size();
        Builder b;
        b.boxes.resize(n);
        for (size_t i = 0; i < n; ++i) b.boxes[i] = items[i].bounding_box();
        b.order.resize(n);
        std::iota(b.order.begin(), b.order.end(), 0u);
        b.out = &bvh.nodes;
        bvh.nodes.reserve(n > 1 ? n - 1 : 0);
        Aabb box;
        bvh.initial_index = b.split(0, n, 0, 0, &box, &bvh.depth);
        return bvh;
    }

private:
    struct Builder {
        std::vector<Aabb> boxes;
        std::vector<uint32_t> order;  // permutation of item indices; sub-ranges are re-sorted in place
        std::vector<RtBvhNode>* out = nullptr;
        uint32_t split(size_t lo, size_t hi, uint32_t axis, uint32_t level, Aabb* box, uint32_t* max_level) {
            const size_t len = hi - lo;
            if (len == 1) {
                *box = boxes[order[lo]];
                *max_level = level;
                return RT_LEAF_BIT | order[lo];
            }
            std::stable_sort(order.begin() + lo, order.begin() + hi,
                             [&](uint32_t a, uint32_t c) { return boxes[a].min[(int)axis] < boxes[c].min[(int)axis]; });
            const uint32_t id = (uint32_t)out->size();
            out->push_back(RtBvhNode{});
            uint32_t left, right, dl = level, dr = level;
            Aabb bl, br;
            if (len == 2) {
                left = RT_LEAF_BIT | order[lo];
                right = RT_LEAF_BIT | order[lo + 1];
                bl = boxes[order[lo]];
                br = boxes[order[lo + 1]];
            } else {
                const size_t mid = len / 2;
                left = split(lo, lo + mid, (uint32_t)((axis + mid) % 3), level + 1, &bl, &dl);
                right = split(lo + mid, hi, (uint32_t)((axis + (len - mid)) % 3), level + 1, &br, &dr);
            }
            const Aabb u = Aabb::surrounding(bl, br);
            RtBvhNode& nd = (*out)[id];
            nd.aabb_min[0] = u.min.x; nd.aabb_min[1] = u.min.y; nd.aabb_min[2] = u.min.z;
            nd.aabb_max[0] = u.max.x; nd.aabb_max[1] = u.max.y; nd.aabb_max[2] = u.max.z;
            nd.left = left; nd.right = right; nd.axis_id = axis;
            *box = u;
            *max_level = std::max(dl, dr);
            return id;
        }
    };
};

// ------------------------------------------------------------------ camera (src/lib/camera.rs)
class Camera {
public:
    Point3 position;
    Dir3 upper_left_corner, unit_right, unit_up, scaled_right, scaled_up;
    float lens_radius = 0.0f;
    Range time_interval{0.0f, 0.0f};
    Vec2f shutter_pace;
    float aspect_ratio() const { return length(scaled_up) / length(scaled_right); }  // :171-173

    class Builder {  // type-state builder of the reference collapsed into one class; same call names
    public:
        Builder& viewport(float w, float h) { vw_ = w; vh_ = h; return *this; }
        Builder& vertical_fov(float fov_deg, float aspect_ratio) {  // :72-79 (sets the viewport WIDTH, Q10)
            const float h = libm_tan((fov_deg * (3.14159265358979323846f / 180.0f)) * 0.5f);
            return viewport(2.0f * h, 2.0f * h * aspect_ratio);
        }
        Builder& position(Point3 p) { pos_ = p; return *this; }
        Builder& orientation(Dir3 up, Dir3 forward) { up_ = up; fwd_ = forward; return *this; }
        Builder& look_at(Dir3 up, Point3 at) { up_ = up; fwd_ = at - pos_; return *this; }
        Builder& look_at_focus(Dir3 up, Point3 at) { look_at(up, at); focus_ = length(fwd_); return *this; }
        Builder& focus_point(Point3 p) { focus_ = length(pos_ - p); return *this; }
        Builder& aperture(float a) { aperture_ = a; return *this; }
        Builder& focus_distance(float d) { focus_ = d; return *this; }
        Builder& motion_blur(float a, float b) { t0_ = a; t1_ = b; return *this; }
        Camera build() const {  // :131-152
            Camera c;
            c.unit_right = unit(cross(fwd_, up_));
            c.unit_up = unit(cross(c.unit_right, fwd_));
            const Dir3 sforward = with_length(fwd_, focus_);
            c.upper_left_corner = focus_ * (c.unit_right * (vw_ * -0.5f) + c.unit_up * (vh_ * 0.5f)) + sforward;
            c.position = pos_;
            c.scaled_right = c.unit_right * (focus_ * vw_);
            c.scaled_up = c.unit_up * (focus_ * vh_);
            c.lens_radius = aperture_ / 2.0f;
            c.time_interval = Range{t0_, t1_};
            return c;
        }

    private:
        float vw_ = 0, vh_ = 0, aperture_ = 0, focus_ = 1.0f, t0_ = 0, t1_ = 0;
        Point3 pos_;
        Dir3 up_ = dir::UP, fwd_ = dir::FORWARD;
    };
    static Builder build() { return Builder(); }  // Camera::build()  :167-169
};

struct BackgroundColor {  // src/lib/background_color.rs
    bool is_sky = true;
    Color color;
    static BackgroundColor Sky() { return {true, {}}; }
    static BackgroundColor Solid(Color c) { return {false, c}; }
};

struct WorldScatteringDistributionProvider {  // ::Rect(RectGeometry)  world_scattering_distribution.rs:4-9
    RectGeometry rect;
};

enum class RenderMode : uint32_t { Default = RT_MODE_DEFAULT, Normals = RT_MODE_NORMALS };

struct HitInteraction {  // hittable.rs:68-76 (+ the flattened element index)
    uint32_t item = RT_MISS;
    float t = 0;
    Dir3 normal;
    Vec2f uv;
    bool front_face = false;
};

[[noreturn]] inline void throw_rt(const char* what) { throw std::runtime_error(std::string(what) + ": " + rt_last_error()); }

// Scene = the flat element list + its BVH (what `Scene::new(BoundingVolumeHierarchy::new(all_elements))` holds).
class Scene {
public:
    std::vector<SurfaceGeometry> elements;  // order == hit id (world_builder.rs:316)
    BoundingVolumeHierarchy bvh;
    explicit Scene(std::vector<SurfaceGeometry> elems) : elements(std::move(elems)), bvh(BoundingVolumeHierarchy::build(elements)) {}
};

// World (rendering.rs:12-17) + the device-resident copy of it.
class World {
public:
    Camera camera;
    std::shared_ptr<Scene> hittable;
    BackgroundColor background;
    bool has_scattering_distribution_provider = false;
    WorldScatteringDistributionProvider scattering_distribution_provider;
    std::shared_ptr<MaterialTable> tables;

    ~World() { release_device(); }
    World() = default;
    World(const World&) = delete;
    World& operator=(const World&) = delete;

    // Flatten into the C-ABI description (pointers stay valid while *this and `items_scratch` live).
    // Uninitialised item records, written on all cores (every field of every record is set below): the flat form of 10^6 elements
    // is 140 MB, and this step is part of every scene hand-off.
    struct ItemArray {
        std::unique_ptr<RtItem[]> p;
        size_t n = 0;
    };
    RtSceneDesc describe(ItemArray& items_scratch, std::vector<RtImage>& images_scratch, std::vector<RtPerlin>& perlins_scratch) const {
        const Scene& sc = *hittable;
        const size_t n_items = sc.elements.size();
        items_scratch.p.reset(new RtItem[n_items]);
        items_scratch.n = n_items;
        RtItem* const items_out = items_scratch.p.get();
        auto fill = [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; ++i) {
            const SurfaceGeometry& e = sc.elements[i];
            RtItem& it = items_out[i];
            e.geometry.write(it);
            it.material = e.material;
            it.wrap_flags = (e.has_transformation ? (uint32_t)RT_WRAP_TRANSFORMATION : 0u) | (e.has_animation ? (uint32_t)RT_WRAP_ANIMATION : 0u) |
                            (e.is_volume ? (uint32_t)RT_WRAP_VOLUME : 0u);
            it.neg_inv_density = e.neg_inv_density;
            it.transform[0] = e.transformation.offset.x; it.transform[1] = e.transformation.offset.y; it.transform[2] = e.transformation.offset.z;
            it.transform[3] = e.transformation.y_sine; it.transform[4] = e.transformation.y_cosine;
            it.velocity[0] = e.velocity.x; it.velocity[1] = e.velocity.y; it.velocity[2] = e.velocity.z;
        }
        };
        {
            const size_t parts = std::max<size_t>(1, std::min<size_t>(std::max(1u, std::thread::hardware_concurrency()), n_items / 32768));
            std::vector<std::thread> pool;
            for (size_t q = 1; q < parts; ++q) pool.emplace_back([&, q] { fill(n_items * q / parts, n_items * (q + 1) / parts); });
            fill(0, n_items / parts);
            for (auto& t : pool) t.join();
        }
        images_scratch.clear();
        for (const ImageData& im : tables->images) images_scratch.push_back(RtImage{im.width, im.height, im.rgb8.data()});
        perlins_scratch.clear();
        for (const PerlinData& pn : tables->perlins) perlins_scratch.push_back(RtPerlin{pn.ranvec.data(), pn.perm_x.data(), pn.perm_y.data(), pn.perm_z.data()});
        RtSceneDesc d{};
        d.n_perlins = (uint32_t)perlins_scratch.size();
        d.perlins = perlins_scratch.data();
        d.n_items = (uint32_t)items_scratch.n;
        d.items = items_scratch.p.get();
        d.n_nodes = (uint32_t)sc.bvh.nodes.size();
        d.nodes = sc.bvh.nodes.data();
        d.root = sc.bvh.initial_index;
        d.n_materials = (uint32_t)tables->materials.size();
        d.materials = tables->materials.data();
        d.n_textures = (uint32_t)tables->textures.size();
        d.textures = tables->textures.data();
        d.n_images = (uint32_t)images_scratch.size();
        d.images = images_scratch.data();
        auto put = [](float* o, Vec3 v) { o[0] = v.x; o[1] = v.y; o[2] = v.z; };
        put(d.camera.position, camera.position);
        put(d.camera.upper_left_corner, camera.upper_left_corner);
        put(d.camera.unit_right, camera.unit_right);
        put(d.camera.unit_up, camera.unit_up);
        put(d.camera.scaled_right, camera.scaled_right);
        put(d.camera.scaled_up, camera.scaled_up);
        d.camera.lens_radius = camera.lens_radius;
        d.camera.time_start = camera.time_interval.start;
        d.camera.time_end = camera.time_interval.end;
        d.camera.shutter_pace[0] = camera.shutter_pace.x;
        d.camera.shutter_pace[1] = camera.shutter_pace.y;
        d.background_is_sky = background.is_sky ? 1u : 0u;
        put(d.background_rgb, background.color);
        d.has_light = has_scattering_distribution_provider ? 1u : 0u;
        const RectGeometry& lr = scattering_distribution_provider.rect;
        d.light_plane = (uint32_t)lr.rect_plane;
        d.light_dist = lr.dist;
        d.light_r0[0] = lr.r0[0]; d.light_r0[1] = lr.r0[1];
        d.light_r1[0] = lr.r1[0]; d.light_r1[1] = lr.r1[1];
        return d;
    }
    // Upload (once) and return the device scene.  Fails loudly without a CUDA device.
    RtScene* device(int device_ordinal = 0) const {
        if (dev_ && dev_ordinal_ == device_ordinal) return dev_;
        release_device();
        ItemArray items;
        std::vector<RtImage> images;
        std::vector<RtPerlin> perlins;
        const auto t_describe = std::chrono::steady_clock::now();
        RtSceneDesc d = describe(items, images, perlins);
        if (std::getenv("RT_PACK_LOG"))
            std::fprintf(stderr, "[host mirror] flat form of the scene %8.1f ms\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_describe).count());
        if (rt_scene_create(&d, device_ordinal, &dev_) != RT_OK) throw_rt("rt_scene_create");
        dev_ordinal_ = device_ordinal;
        return dev_;
    }
    void release_device() const {
        if (dev_) rt_scene_destroy(dev_);
        dev_ = nullptr;
    }
    // Scene::hit (hittable.rs:130-137) for a batch of rays.
    std::vector<HitInteraction> hit(const std::vector<Ray>& rays, Range t_range, int device_ordinal = 0) const {
        const size_t n = rays.size();
        std::vector<float> org(3 * n), dirs(3 * n), t(n), nrm(3 * n), uv(2 * n);
        std::vector<uint32_t> id(n);
        std::vector<uint8_t> ff(n);
        for (size_t i = 0; i < n; ++i) {
            org[3 * i] = rays[i].origin.x; org[3 * i + 1] = rays[i].origin.y; org[3 * i + 2] = rays[i].origin.z;
            dirs[3 * i] = rays[i].direction.x; dirs[3 * i + 1] = rays[i].direction.y; dirs[3 * i + 2] = rays[i].direction.z;
        }
        if (rt_trace(device(device_ordinal), org.data(), dirs.data(), n, t_range.start, t_range.end, id.data(), t.data(),
                     nrm.data(), ff.data(), uv.data()) != RT_OK)
            throw_rt("rt_trace");
        std::vector<HitInteraction> out(n);
        for (size_t i = 0; i < n; ++i) {
            out[i].item = id[i]; out[i].t = t[i];
            out[i].normal = {nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]};
            out[i].uv = {uv[2 * i], uv[2 * i + 1]};
            out[i].front_face = ff[i] != 0;
        }
        return out;
    }

private:
    mutable RtScene* dev_ = nullptr;
    mutable int dev_ordinal_ = -1;
};

// render() — src/lib/rendering.rs:121-128.  Same arguments and result layout (W*H linear-RGB colours, row 0 = top);
// `thread_count` keeps its role in the sample split (split_work_tasks, :222-237) and becomes the number of sample
// lanes in flight per pixel on the GPU.  `seed` replaces the reference's OS-entropy seeding (:160).
// render_into: the same call writing into a caller-owned buffer of image_size.count() colours (no allocation, no extra copy).
inline void render_into(Size2i image_size, size_t thread_count, size_t samples_per_pixel, int32_t max_depth, const World& world,
                        RenderMode render_mode, Color* out, uint64_t seed = 0x5EEDull, RtStats* stats = nullptr, int device_ordinal = 0) {
    RtRenderParams p{};
    p.width = image_size.width; p.height = image_size.height;
    p.samples_per_pixel = samples_per_pixel;
    p.max_depth = max_depth;
    p.thread_count = (uint32_t)thread_count;
    p.render_mode = (uint32_t)render_mode;
    p.seed = seed;
    static_assert(sizeof(Color) == 12, "Color must be 3 packed floats");
    if (rt_render(world.device(device_ordinal), &p, reinterpret_cast<float*>(out), stats) != RT_OK) throw_rt("rt_render");
}
inline std::vector<Color> render(Size2i image_size, size_t thread_count, size_t samples_per_pixel, int32_t max_depth,
                                 const World& world, RenderMode render_mode, uint64_t seed = 0x5EEDull, RtStats* stats = nullptr,
                                 int device_ordinal = 0) {
    std::vector<Color> out(image_size.count());
    render_into(image_size, thread_count, samples_per_pixel, max_depth, world, render_mode, out.data(), seed, stats, device_ordinal);
    return out;
}

// Color::to_rgb8_gamma2 (color.rs:43-48) — host-side, for callers that save images (src/app/main.rs:53-56)
inline void to_rgb8_gamma2(Color c, uint8_t out[3]) {
    const float v[3] = {c.x, c.y, c.z};
    for (int i = 0; i < 3; ++i) {
        float s = std::sqrt(v[i]) * 256.0f;
        float cl = s > 255.0f ? 255.0f : (s < 0.0f ? 0.0f : s);
        out[i] = (cl == cl) ? (uint8_t)cl : (uint8_t)0;
    }
}

}  // namespace rtiaw
