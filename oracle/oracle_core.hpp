// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path: only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may use it.
// PARITY UNPINNED: the reference (martingleich/RayTracingInAWeekend) ships no tests, golden
// vectors or fixtures for this path and cannot be compiled in this image (no rustc/cargo), so
// this restatement is checked only against analytic known-answer tests (tests/test_oracle_kat.py).
//
// Literal CPU restatement (C++17, compile with -O2 -ffp-contract=off -fno-fast-math) of the
// reference's `src/lib` crate — quirks included.  Each block cites the
// reference file:line it follows.  Arithmetic is single-rounded f32 in the reference's operand
// order (Rust never contracts a*b+c).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <deque>
#include <limits>
#include <memory>
#include <vector>

#include "oracle_rng.hpp"

namespace orc {

static const float PI_F = 3.14159265358979323846f;
static const float TAU_F = 6.28318530717958647692f;
static const float INF_F = std::numeric_limits<float>::infinity();
// libm calls the reference makes at RUN time (f32::tan / sin_cos -> the platform libm).  `volatile` keeps the compiler from
// folding them at compile time with a differently-rounded (MPFR) result.
static inline float rt_tanf(float x) { volatile float v = x; return std::tan(v); }
static inline float rt_sinf(float x) { volatile float v = x; return std::sin(v); }
static inline float rt_cosf(float x) { volatile float v = x; return std::cos(v); }

// ---------------------------------------------------------------- src/lib/vec3.rs:1-306
struct V3 {
    float e[3];
    V3() : e{0, 0, 0} {}
    V3(float a, float b, float c) : e{a, b, c} {}
};
static inline V3 operator+(V3 a, V3 b) { return V3(a.e[0] + b.e[0], a.e[1] + b.e[1], a.e[2] + b.e[2]); }
static inline V3 operator-(V3 a, V3 b) { return V3(a.e[0] - b.e[0], a.e[1] - b.e[1], a.e[2] - b.e[2]); }
static inline V3 operator-(V3 a) { return V3(-a.e[0], -a.e[1], -a.e[2]); }
static inline V3 operator*(V3 a, float s) { return V3(a.e[0] * s, a.e[1] * s, a.e[2] * s); }
static inline V3 operator*(float s, V3 a) { return a * s; }  // vec3.rs:299-305 (rhs * self)
static inline V3 operator/(V3 a, float s) { return V3(a.e[0] / s, a.e[1] / s, a.e[2] / s); }
static inline bool operator==(V3 a, V3 b) { return a.e[0] == b.e[0] && a.e[1] == b.e[1] && a.e[2] == b.e[2]; }
static inline bool operator!=(V3 a, V3 b) { return !(a == b); }
typedef V3 Point3;
typedef V3 Dir3;
static const Dir3 DIR_UP(0.0f, 1.0f, 0.0f);

static inline float dot(V3 a, V3 b) { return a.e[0] * b.e[0] + a.e[1] * b.e[1] + a.e[2] * b.e[2]; }  // :189-191
static inline V3 cross(V3 a, V3 b) {                                                                  // :193-201
    return V3(a.e[1] * b.e[2] - a.e[2] * b.e[1], a.e[2] * b.e[0] - a.e[0] * b.e[2],
              a.e[0] * b.e[1] - a.e[1] * b.e[0]);
}
static inline float length_squared(V3 a) { return dot(a, a); }
static inline float length(V3 a) { return std::sqrt(length_squared(a)); }
static inline V3 with_length(V3 a, float len) { return a * (len / length(a)); }  // :206-208
static inline V3 unit(V3 a) { return with_length(a, 1.0f); }                     // :203-205
static inline V3 unit_or_else(V3 a, V3 dflt) {                                   // :223-230
    float len_sq = length_squared(a);
    if (len_sq > 1e-8f) return a / std::sqrt(len_sq);
    return dflt;
}
static inline V3 reflect(V3 ray, V3 normal) { return ray - (2.0f * dot(ray, normal)) * normal; }  // :232-234
static inline V3 refract(V3 ray, V3 normal, float etai_over_etat) {                               // :235-240
    float cos_theta = std::fmin(dot(-ray, normal), 1.0f);
    V3 r_out_perp = etai_over_etat * (ray + cos_theta * normal);
    V3 r_out_parallel = -std::sqrt(std::fabs(1.0f - length_squared(r_out_perp))) * normal;
    return r_out_perp + r_out_parallel;
}

// ---------------------------------------------------------------- src/lib/vec2.rs
struct Vec2f {
    float x, y;
    Vec2f() : x(0), y(0) {}
    Vec2f(float a, float b) : x(a), y(b) {}
};
static inline Vec2f operator+(Vec2f a, Vec2f b) { return Vec2f(a.x + b.x, a.y + b.y); }
static inline Vec2f operator*(Vec2f a, float s) { return Vec2f(a.x * s, a.y * s); }

// ---------------------------------------------------------------- src/lib/math.rs
template <class T>
static inline T lerp(T a, T b, float t) { return a * (1.0f - t) + b * t; }  // :1-7
template <class T>
static inline T interpolate(float w0, float w1, float w2, const T v[3]) {   // :9-16
    return v[0] * w0 + v[1] * w1 + v[2] * w2;
}
static inline float clampf(float low, float high, float value) {            // :18-27
    if (value > high) return high;
    if (value < low) return low;
    return value;
}

// ---------------------------------------------------------------- src/lib/color.rs
typedef V3 Color;
static inline Color convolution(Color a, Color b) { return Color(a.e[0] * b.e[0], a.e[1] * b.e[1], a.e[2] * b.e[2]); }
static inline Color new_rgb8(const uint8_t c[3]) {  // :29-35
    return Color((float)c[0] / 255.0f, (float)c[1] / 255.0f, (float)c[2] / 255.0f);
}
static inline uint8_t sat_u8(float v) {  // Rust `as u8`: saturating, NaN -> 0
    if (!(v == v)) return 0;
    if (v <= 0.0f) return 0;
    if (v >= 255.0f) return 255;
    return (uint8_t)v;
}
static inline void to_rgb8_gamma2(Color c, uint8_t out[3]) {  // :43-48
    for (int i = 0; i < 3; ++i) out[i] = sat_u8(clampf(0.0f, 255.0f, std::sqrt(c.e[i]) * 256.0f));
}

// ---------------------------------------------------------------- src/lib/ray.rs
struct Ray {
    Point3 origin;
    Dir3 direction;
    float time;
    Point3 at(float t) const { return origin + t * direction; }  // :20-22
};
struct Range {
    float start, end;
    bool contains(float t) const { return start <= t && t < end; }  // half-open, NaN -> false
};

// ---------------------------------------------------------------- src/lib/aabb.rs:10-78
struct Aabb {
    Point3 min, max;
};
static inline Aabb aabb_surrounding(const Aabb& a, const Aabb& b) {  // :26-37 (f32::min/max ignore NaN)
    Aabb r;
    for (int i = 0; i < 3; ++i) {
        r.min.e[i] = std::fmin(a.min.e[i], b.min.e[i]);
        r.max.e[i] = std::fmax(a.max.e[i], b.max.e[i]);
    }
    return r;
}
static inline Aabb aabb_surrounding_points(const Point3* p, int n) {  // :52-63
    Aabb r;
    r.min = p[0];
    r.max = p[0];
    for (int k = 1; k < n; ++k)
        for (int i = 0; i < 3; ++i) {
            r.min.e[i] = std::fmin(r.min.e[i], p[k].e[i]);
            r.max.e[i] = std::fmax(r.max.e[i], p[k].e[i]);
        }
    return r;
}
// Q1: per-axis-independent overlap test, NOT a slab test (t_min/t_max are loop-local).
static inline bool aabb_hit_cond(const Aabb& b, const Ray& ray, const Range& t_range) {  // :65-78
    for (int i = 0; i < 3; ++i) {
        float a0 = (b.min.e[i] - ray.origin.e[i]) / ray.direction.e[i];
        float a1 = (b.max.e[i] - ray.origin.e[i]) / ray.direction.e[i];
        float t0, t1;
        if (a0 < a1) { t0 = a0; t1 = a1; } else { t0 = a1; t1 = a0; }  // math.rs:35-41 minmax
        float t_min = std::fmax(t0, t_range.start);
        float t_max = std::fmin(t1, t_range.end);
        if (t_max <= t_min) return false;
    }
    return true;
}

// ---------------------------------------------------------------- src/lib/hittable.rs:24-54
struct GeoHit {
    Point3 position;
    Dir3 normal;
    Vec2f uv;
    float t;
    bool front_face;
};
static inline GeoHit geo_hit_from_ray(const Ray& ray, Point3 position, Dir3 surface_normal, float t, Vec2f uv) {
    GeoHit g;
    g.front_face = dot(surface_normal, ray.direction) < 0.0f;  // :41
    g.normal = g.front_face ? surface_normal : -surface_normal;
    g.position = position;
    g.t = t;
    g.uv = uv;
    return g;
}

// ---------------------------------------------------------------- src/lib/aabb.rs:80-192 (AABB as a primitive; row f2)
static inline float signum_f32(float v) {  // f32::signum: 1.0 for +0.0 and positives, -1.0 for -0.0 and negatives, NaN for NaN
    if (v != v) return v;
    return std::signbit(v) ? -1.0f : 1.0f;
}
static inline bool aabb_intersections_line(const Aabb& b, Point3 origin, Dir3 direction, float* near_t, int* near_plane,
                                           float* far_t, int* far_plane) {  // :103-167
    Dir3 mn = b.min - origin, mx = b.max - origin;
    float near = -INF_F, far = INF_F;
    int np = 0, fp = 0;
    for (int i = 0; i < 3; ++i) {
        float t1 = mn.e[i] / direction.e[i];
        float t2 = mx.e[i] / direction.e[i];
        float t_min = std::fmin(t1, t2);
        float t_max = std::fmax(t1, t2);
        if (t_min > near) { near = t_min; np = i; }
        if (t_max < far) { far = t_max; fp = i; }
        if (near > far || far < 0.0f) return false;
    }
    *near_t = near; *near_plane = np; *far_t = far; *far_plane = fp;
    return true;
}
static inline bool aabb_prim_hit(const Aabb& b, const Ray& ray, const Range& t_range, GeoHit* out) {  // :80-102
    float near_t, far_t;
    int near_plane, far_plane;
    if (!aabb_intersections_line(b, ray.origin, ray.direction, &near_t, &near_plane, &far_t, &far_plane)) return false;
    float t;
    int plane;
    if (t_range.contains(near_t)) { t = near_t; plane = near_plane; }
    else if (t_range.contains(far_t)) { t = far_t; plane = far_plane; }
    else return false;
    Point3 position = ray.origin + t * ray.direction;
    float center = (b.max.e[plane] + b.min.e[plane]) * 0.5f;
    Dir3 surface_normal(0, 0, 0);
    surface_normal.e[plane] = signum_f32(position.e[plane] - center);
    *out = geo_hit_from_ray(ray, position, surface_normal, t, Vec2f(0, 0));
    return true;
}
// Aabb::corners (:178-192) + Transformation::apply_aabb (hittable.rs:285-291).  Q2: the loop mutates COPIES of the corners
// (`for mut c in corners`), so the result is the bounding box of the UNTRANSFORMED corners — which can still differ from
// the input by one rounding (min + (max - min) != max).
static inline Aabb aabb_apply_transform_q2(const Aabb& a) {
    Dir3 right(a.max.e[0] - a.min.e[0], 0.0f, 0.0f), up(0.0f, a.max.e[1] - a.min.e[1], 0.0f), fwd(0.0f, 0.0f, a.max.e[2] - a.min.e[2]);
    Point3 c[8] = {a.min, a.min + right, a.min + up, a.min + fwd, a.max, a.max - right, a.max - up, a.max - fwd};
    return aabb_surrounding_points(c, 8);
}

// ---------------------------------------------------------------- src/lib/hittable/sphere_geometry.rs
struct SphereGeometry {
    Point3 center;
    float radius;
};
static inline Aabb sphere_bounding_box(const SphereGeometry& s) {  // :17-19, aabb.rs:20-25
    Aabb b;
    b.min = s.center - Dir3(s.radius, s.radius, s.radius);
    b.max = s.center + Dir3(s.radius, s.radius, s.radius);
    return b;
}
static inline Vec2f sphere_uv(Dir3 pos) {  // :56-59 + vec3.rs:241-249 (to_radian)
    float theta = std::acos(pos.e[1]);
    float phi = std::atan2(-pos.e[2], pos.e[0]) + PI_F;
    return Vec2f(phi / TAU_F, theta / PI_F);
}
static inline bool sphere_hit(const SphereGeometry& s, const Ray& ray, const Range& t_range, GeoHit* out) {  // :21-54
    Dir3 oc = ray.origin - s.center;
    float half_b = dot(oc, ray.direction);
    float c = length_squared(oc) - s.radius * s.radius;
    float disc = half_b * half_b - c;
    if (disc < 0.0f) return false;
    float sqrtd = std::sqrt(disc);
    float root_small = -half_b - sqrtd;
    float t;
    if (t_range.contains(root_small)) {
        t = root_small;
    } else {
        float root_large = -half_b + sqrtd;
        if (t_range.contains(root_large)) t = root_large;
        else return false;
    }
    Point3 position = ray.at(t);
    Dir3 surface_normal = (position - s.center) / s.radius;
    *out = geo_hit_from_ray(ray, position, surface_normal, t, sphere_uv(surface_normal));
    return true;
}

// ---------------------------------------------------------------- src/lib/hittable/rect_geometry.rs
enum RectPlane { RECT_XY = 0, RECT_XZ = 1, RECT_YZ = 2 };
static inline void rect_axes(int plane, int* p0, int* p1, int* n) {  // :14-21
    if (plane == RECT_XY) { *p0 = 0; *p1 = 1; *n = 2; }
    else if (plane == RECT_XZ) { *p0 = 0; *p1 = 2; *n = 1; }
    else { *p0 = 1; *p1 = 2; *n = 0; }
}
struct RectGeometry {
    int rect_plane;
    float dist;
    float r0[2], r1[2];
};
static inline bool rect_hit(const RectGeometry& g, const Ray& ray, const Range& t_range, GeoHit* out) {  // :33-59
    int p0, p1, n;
    rect_axes(g.rect_plane, &p0, &p1, &n);
    float t = (g.dist - ray.origin.e[n]) / ray.direction.e[n];
    if (t_range.contains(t)) {
        Point3 position = ray.origin + t * ray.direction;
        if (position.e[p0] >= g.r0[0] && position.e[p0] <= g.r0[1] && position.e[p1] >= g.r1[0] &&
            position.e[p1] <= g.r1[1]) {
            // Q5: uv.y divides by (r0.1 - r1.0)
            Vec2f uv((position.e[p0] - g.r0[0]) / (g.r0[1] - g.r0[0]), (position.e[p1] - g.r1[0]) / (g.r0[1] - g.r1[0]));
            Dir3 surface_normal(0, 0, 0);
            surface_normal.e[n] = -1.0f;
            *out = geo_hit_from_ray(ray, position, surface_normal, t, uv);
            return true;
        }
    }
    return false;
}
static inline Dir3 rect_generate(const RectGeometry& g, Point3 origin, TRng& rng) {  // :60-67
    int p0, p1, n;
    rect_axes(g.rect_plane, &p0, &p1, &n);
    float e[3] = {0, 0, 0};
    e[p0] = rng.gen_range(g.r0[0], g.r0[1]);
    e[p1] = rng.gen_range(g.r1[0], g.r1[1]);
    e[n] = g.dist;
    return unit(Point3(e[0], e[1], e[2]) - origin);
}
static inline float rect_value(const RectGeometry& g, Point3 origin, Dir3 direction) {  // :68-85
    Ray r{origin, direction, 0.0f};
    Range tr{0.001f, INF_F};
    GeoHit hi;
    if (rect_hit(g, r, tr, &hi)) {
        float area = (g.r0[1] - g.r0[0]) * (g.r1[1] - g.r1[0]);
        float distance_squared = hi.t * hi.t;
        float cosine = std::fabs(dot(hi.normal, direction));
        return distance_squared / (cosine * area);
    }
    return 0.0f;
}
static inline Aabb rect_bounding_box(const RectGeometry& g, float thickness) {  // :87-98
    int p0, p1, n;
    rect_axes(g.rect_plane, &p0, &p1, &n);
    Aabb b;
    b.min.e[p0] = g.r0[0]; b.min.e[p1] = g.r1[0]; b.min.e[n] = g.dist - thickness;
    b.max.e[p0] = g.r0[1]; b.max.e[p1] = g.r1[1]; b.max.e[n] = g.dist + thickness;
    return b;
}

// ---------------------------------------------------------------- src/lib/hittable/triangle_geometry.rs
struct TriangleGeometry {
    Point3 positions[3];
    Dir3 normals[3];
    Vec2f texture_coords[3];
};
static inline bool triangle_hit(const TriangleGeometry& g, const Ray& ray, const Range& t_range, GeoHit* out) {  // :13-45
    Point3 p0 = g.positions[0], p1 = g.positions[1], p2 = g.positions[2];
    Dir3 dir1 = p1 - p0;
    Dir3 dir2 = p2 - p0;
    Dir3 normal = unit(cross(dir1, dir2));
    float denom = dot(ray.direction, normal);
    if (std::fabs(denom) > 0.0001f) {
        float t = dot(p0 - ray.origin, normal) / denom;
        if (t_range.contains(t)) {
            Point3 position = ray.at(t);
            Dir3 q = position - p0;
            Dir3 v_temp = cross(normal, dir2);
            float w1 = dot(q, v_temp) / dot(dir1, v_temp);
            if (w1 > 0.0f && w1 < 1.0f) {
                Dir3 v_temp2 = cross(normal, dir1);
                float w2 = dot(q, v_temp2) / dot(dir2, v_temp2);
                float w0 = 1.0f - w1 - w2;
                if (w2 > 0.0f && w0 > 0.0f) {
                    Vec2f uv = interpolate(w0, w1, w2, g.texture_coords);
                    Dir3 surface_normal = interpolate(w0, w1, w2, g.normals);  // Q8: not re-normalised
                    *out = geo_hit_from_ray(ray, position, surface_normal, t, uv);
                    return true;
                }
            }
        }
    }
    return false;
}
static inline Aabb triangle_bounding_box(const TriangleGeometry& g) { return aabb_surrounding_points(g.positions, 3); }  // :47-49

// ---------------------------------------------------------------- src/lib/transformations.rs
static inline void rotate_around_up(float c, float s, V3& v) {  // :107-111
    float x = v.e[0], y = v.e[2];
    v.e[0] = c * x + s * y;
    v.e[2] = -s * x + c * y;
}
struct Transformation {
    Dir3 offset;
    float y_sine = 0.0f, y_cosine = 1.0f;
    bool is_zero() const { return offset == Dir3(0, 0, 0) && y_sine == 0.0f && y_cosine == 1.0f; }
    Transformation translate(Dir3 o) const {  // :18-22
        Transformation r = *this;
        r.offset = r.offset + o;
        return r;
    }
    Transformation rotate_around_up_deg(float angle) const {  // :26-36
        Transformation r = *this;
        float rad = angle * (PI_F / 180.0f);  // f32::to_radians
        float s = rt_sinf(rad), c = rt_cosf(rad);
        rotate_around_up(c, s, r.offset);
        V3 v(y_sine, 0.0f, y_cosine);
        rotate_around_up(c, s, v);
        r.y_sine = v.e[0];
        r.y_cosine = v.e[2];
        return r;
    }
    Point3 apply_point(Point3 p) const {  // :37-40
        rotate_around_up(y_cosine, y_sine, p);
        return p + offset;
    }
    Dir3 apply_normal(Dir3 d) const {  // :42-44
        rotate_around_up(y_cosine, y_sine, d);
        return d;
    }
    Dir3 apply_direction(Dir3 d) const { return apply_normal(d); }
    Point3 reverse_point(Point3 p) const {  // :49-52
        p = p - offset;
        rotate_around_up(y_cosine, -y_sine, p);
        return p;
    }
    Dir3 reverse_normal(Dir3 d) const {  // :54-56
        rotate_around_up(y_cosine, -y_sine, d);
        return d;
    }
    Transformation then(const Transformation& next) const {  // :89-98
        Transformation r;
        r.offset = next.apply_direction(offset) + next.offset;
        r.y_sine = y_sine * next.y_cosine + y_cosine * next.y_sine;
        r.y_cosine = y_cosine * next.y_cosine - y_sine * next.y_sine;
        return r;
    }
};

// ---------------------------------------------------------------- src/lib/perlin.rs
struct Perlin {
    std::vector<Dir3> ranvec;
    std::vector<uint32_t> perm_x, perm_y, perm_z;
    size_t bits = 0;
    static Perlin make(size_t bits, TRng& rng) {  // :19-36  (scene-construction RNG: FAITHFUL stream)
        Perlin p;
        p.bits = bits;
        const size_t n = (size_t)1 << bits;
        for (size_t i = 0; i < n; ++i) { float s[3]; rng.unit_sphere(s); p.ranvec.push_back(Dir3(s[0], s[1], s[2])); }
        auto gen = [&](std::vector<uint32_t>& v) {  // generate_per :93-97
            v.resize(n);
            for (size_t i = 0; i < n; ++i) v[i] = (uint32_t)i;
            rng.shuffle(v.data(), n);
        };
        gen(p.perm_x); gen(p.perm_y); gen(p.perm_z);
        return p;
    }
    float noise(Point3 p) const {  // :48-91
        const int mask = (int)(((size_t)1 << bits) - 1);
        float fx = std::floor(p.e[0]), fy = std::floor(p.e[1]), fz = std::floor(p.e[2]);
        float u = p.e[0] - fx, v = p.e[1] - fy, w = p.e[2] - fz;
        int i = (int)fx, j = (int)fy, k = (int)fz;
        float uu = u * u * (3.0f - 2.0f * u), vv = v * v * (3.0f - 2.0f * v), ww = w * w * (3.0f - 2.0f * w);
        float accum = 0.0f;
        for (int di = 0; di < 2; ++di)
            for (int dj = 0; dj < 2; ++dj)
                for (int dk = 0; dk < 2; ++dk) {
                    uint32_t r = perm_x[(size_t)((i + di) & mask)] ^ perm_y[(size_t)((j + dj) & mask)] ^ perm_z[(size_t)((k + dk) & mask)];
                    Dir3 c = ranvec[r];
                    float fi = (float)di, fj = (float)dj, fk = (float)dk;
                    Dir3 weight(u - (float)di, v - (float)dj, w - (float)dk);
                    accum += (fi * uu + (1.0f - fi) * (1.0f - uu)) * (fj * vv + (1.0f - fj) * (1.0f - vv)) *
                             (fk * ww + (1.0f - fk) * (1.0f - ww)) * dot(c, weight);
                }
        return accum;
    }
    float turbulence(Point3 p, int depth, float fall_off) const {  // :37-47
        float accum = 0.0f, weight = 1.0f;
        for (int d = 0; d < depth; ++d) {
            accum += weight * noise(p);
            weight *= fall_off;
            p = p * 2.0f;
        }
        return std::fabs(accum);
    }
};

// ---------------------------------------------------------------- src/lib/texture.rs + material.rs
struct Image {
    uint32_t width = 0, height = 0;
    std::vector<uint8_t> rgb;  // row 0 = top
};
struct HitInteraction;
struct Texture {
    enum Kind { Solid = 0, Image_ = 1, Checker = 2, Marble = 3 } kind = Solid;
    float scale = 0.0f;                      // Marble (texture.rs:13-16)
    const Perlin* noise = nullptr;
    Color color;
    const Image* image = nullptr;
    float inv_frequency = 0.0f;              // Checker (texture.rs:8-12)
    const Texture *even = nullptr, *odd = nullptr;
    Color sample(const HitInteraction& hi) const;
};
struct Material {
    enum Kind { Lambert = 0, Metal = 1, Dielectric = 2, DiffuseLight = 3, Isotropic = 4 } kind = Lambert;
    const Texture* tex = nullptr;  // albedo / emit
    float param = 0.0f;            // fuzz / index_of_refraction
};
struct HitInteraction {  // hittable.rs:68-76 (+ item: the flattened element index = "hit ID", world_builder.rs:316)
    Point3 position;
    Dir3 normal;
    Vec2f uv;
    float t;
    bool front_face;
    const Material* material;
    uint32_t item;
};
static inline uint32_t sat_u32(float v) {  // Rust `as u32`
    if (!(v == v)) return 0;
    if (v <= 0.0f) return 0;
    if (v >= 4294967296.0f) return 0xFFFFFFFFu;
    return (uint32_t)v;
}
inline Color Texture::sample(const HitInteraction& hi) const {  // texture.rs:23-43
    if (kind == Solid) return color;
    if (kind == Marble) {  // :44-51
        float v = 0.5f * (1.0f + std::sin(hi.position.e[2] * scale + 10.0f * noise->turbulence(hi.position, 7, 0.5f)));
        return Color(v, v, v);  // Color(1,1,1) * 0.5 * (1 + sin(..)): 1*0.5 is exact
    }
    if (kind == Checker) {  // :26-35
        Dir3 sp = (hi.position - Point3(0, 0, 0)) * inv_frequency;
        float sines = std::sin(sp.e[0]) * std::sin(sp.e[1]) * std::sin(sp.e[2]);
        return (sines < 0.0f ? even : odd)->sample(hi);
    }
    uint32_t pu = std::min(sat_u32(hi.uv.x * (float)image->width), image->width - 1);
    uint32_t pv = std::min(sat_u32(hi.uv.y * (float)image->height), image->height - 1);
    return new_rgb8(&image->rgb[3 * ((size_t)pv * image->width + pu)]);
}

struct ScatterDist {  // material.rs:7-40
    bool mirror;
    Dir3 v;  // Cosine(normal) or Mirror(direction)
    Dir3 generate(TRng& rng) const {
        if (mirror) return v;
        float s[3];
        rng.unit_sphere(s);
        return unit_or_else(v + Dir3(s[0], s[1], s[2]), v);
    }
    float value(Dir3 direction) const {
        if (mirror) return INF_F;
        return std::fmax(dot(v, direction), 0.0f) / PI_F;
    }
};
static inline float reflectance(float cosine, float ref_idx) {  // material.rs:83-87 (powi(5) = x*(x^2)^2)
    float r0 = (1.0f - ref_idx) / (1.0f + ref_idx);
    float rs = r0 * r0;
    float x = 1.0f - cosine;
    float x2 = x * x;
    float x4 = x2 * x2;
    return rs + (1.0f - rs) * (x * x4);
}
static inline bool material_scatter(const Material& m, const Ray& ray, const HitInteraction& hi, TRng& rng,
                                    Color* att, ScatterDist* dist) {  // material.rs:52-114
    switch (m.kind) {
        case Material::Lambert:
            *att = m.tex->sample(hi);
            *dist = ScatterDist{false, hi.normal};
            return true;
        case Material::Metal: {
            Dir3 fuzz_dir(0, 0, 0);
            if (m.param > 0.0f) {
                float b[3];
                rng.unit_ball(b);
                fuzz_dir = m.param * Dir3(b[0], b[1], b[2]);
            }
            Dir3 direction = reflect(ray.direction, hi.normal) + fuzz_dir;
            if (dot(direction, hi.normal) > 0.0f) {
                *dist = ScatterDist{true, unit(direction)};
                *att = m.tex->sample(hi);
                return true;
            }
            return false;
        }
        case Material::Dielectric: {
            float refraction_ratio = hi.front_face ? 1.0f / m.param : m.param;
            float cos_theta = std::fmin(dot(-ray.direction, hi.normal), 1.0f);
            float sin_theta = std::sqrt(1.0f - cos_theta * cos_theta);
            bool cannot_refract = refraction_ratio * sin_theta > 1.0f;
            Dir3 direction;
            if (cannot_refract || (reflectance(cos_theta, refraction_ratio) > rng.gen_f32()))
                direction = reflect(ray.direction, hi.normal);
            else
                direction = refract(ray.direction, hi.normal, refraction_ratio);
            *dist = ScatterDist{true, unit(direction)};
            *att = Color(1, 1, 1);
            return true;
        }
        case Material::Isotropic: {  // :106-111
            float sp[3];
            rng.unit_sphere(sp);
            *dist = ScatterDist{true, Dir3(sp[0], sp[1], sp[2])};
            *att = m.tex->sample(hi);
            return true;
        }
        default:
            return false;
    }
}
static inline float material_scattering_pdf(const Material& m, const Ray& scattered, const HitInteraction& hi) {  // :116-130
    if (m.kind == Material::Lambert) {
        float cosine = dot(hi.normal, scattered.direction);
        return std::fmax(cosine, 0.0f) / PI_F;
    }
    return 0.0f;
}
static inline Color material_emit(const Material& m, const HitInteraction& hi) {  // :132-137
    if (m.kind == Material::DiffuseLight) return m.tex->sample(hi);
    return Color(0, 0, 0);
}

// ---------------------------------------------------------------- src/lib/hittable.rs:104-482
struct Geometry {
    enum Kind { Sphere = 0, Rect = 1, Triangle = 2, AxisAlignedBox = 3 } kind;
    SphereGeometry sphere;
    RectGeometry rect;
    TriangleGeometry tri;
    Aabb box;
    bool hit(const Ray& ray, const Range& tr, GeoHit* out) const {  // :141-148
        switch (kind) {
            case Sphere: return sphere_hit(sphere, ray, tr, out);
            case Rect: return rect_hit(rect, ray, tr, out);
            case AxisAlignedBox: return aabb_prim_hit(box, ray, tr, out);
            default: return triangle_hit(tri, ray, tr, out);
        }
    }
    Aabb bounding_box() const {  // :150-157
        switch (kind) {
            case Sphere: return sphere_bounding_box(sphere);
            case Rect: return rect_bounding_box(rect, 0.01f);
            case AxisAlignedBox: return box;
            default: return triangle_bounding_box(tri);
        }
    }
    // :165-208; *rest receives the transformation that must stay as a run-time SceneElement::Transformation wrapper
    Geometry partial_apply_transformation(const Transformation& t, bool* has_rest, Transformation* rest) const {
        Geometry g = *this;
        *has_rest = false;
        if (kind == Sphere) {
            g.sphere.center = t.apply_point(sphere.center);
        } else if (kind == Triangle) {
            for (int i = 0; i < 3; ++i) {
                g.tri.positions[i] = t.apply_point(tri.positions[i]);
                g.tri.normals[i] = t.apply_normal(tri.normals[i]);
            }
        } else if (kind == AxisAlignedBox) {  // Q3: the translation is baked, the rotation-only remainder is applied afterwards
            g.box.min = box.min + t.offset;
            g.box.max = box.max + t.offset;
            Transformation rem = t;
            rem.offset = Dir3(0, 0, 0);
            *has_rest = !rem.is_zero();
            *rest = rem;
        } else {
            *has_rest = !t.is_zero();
            *rest = t;
        }
        return g;
    }
};

struct Counters {
    uint64_t node_tests = 0, prim_tests = 0, rays = 0;
};

struct BvhNode {  // :345-351
    Aabb aabb;
    size_t axis_id = 0, left = 0, right = 0;
};
static const size_t USIZE_MAX = ~(size_t)0;

// SceneElement::SurfaceGeometry(geo, material) (:115), optionally wrapped as Transformation(..) (:118) and then Animation(..)
// (:117) exactly in the order NodeRef::finish applies them (world_builder.rs:310-315).
struct SceneItem {
    Geometry geo;
    const Material* material;
    bool has_transform = false;
    Transformation transform;
    bool has_animation = false;
    Dir3 velocity;
    bool is_volume = false;        // SceneElement::VolumeGeometry (:116, 294-342): `material` is the phase function
    float neg_inv_density = 0.0f;  // -1 / density

    // VolumeGeometry::hit :315-342.  Returns t via out->t; normal UP, uv ZERO, front_face false (all "arbitrary" in the reference).
    bool volume_hit(const Ray& ray, const Range& t_range, TRng& rng, uint32_t item, GeoHit* out) const {
        GeoHit b0, b1;
        if (!geo.hit(ray, Range{-INF_F, INF_F}, &b0)) return false;
        float start_boundary = b0.t;
        float next = start_boundary + 0.001f;
        if (!geo.hit(ray, Range{next, INF_F}, &b1)) return false;
        float end_boundary = b1.t;
        float start_medium = std::fmax(start_boundary, t_range.start);
        float end_medium = std::fmin(end_boundary, t_range.end);
        if (start_medium >= end_medium) return false;
        float t = std::fmax(start_medium, 0.0f) + neg_inv_density * rng.volume_ln_u(item);
        if (t > end_medium) return false;
        out->position = ray.at(t);
        out->normal = Dir3(0.0f, 1.0f, 0.0f);
        out->uv = Vec2f(0, 0);
        out->t = t;
        out->front_face = false;
        return true;
    }

    Aabb bounding_box() const {  // SceneElement::bounding_box :249-267 (Q2: wrappers do not move / extend the box)
        Aabb b = geo.bounding_box();
        if (has_transform) b = aabb_apply_transform_q2(b);
        if (has_animation) {
            Aabb s = aabb_apply_transform_q2(b), e = aabb_apply_transform_q2(b);
            b = aabb_surrounding(s, e);
        }
        return b;
    }
    bool hit(const Ray& ray_in, const Range& tr, TRng& rng, uint32_t item, GeoHit* out) const {  // SceneElement::hit :230-244
        Ray ray = ray_in;
        Transformation anim;
        if (has_animation) {
            anim = Transformation().translate(velocity * ray.time);
            ray.origin = anim.reverse_point(ray.origin);
            ray.direction = anim.reverse_normal(ray.direction);
        }
        if (has_transform) {
            ray.origin = transform.reverse_point(ray.origin);
            ray.direction = transform.reverse_normal(ray.direction);
        }
        if (is_volume ? !volume_hit(ray, tr, rng, item, out) : !geo.hit(ray, tr, out)) return false;
        if (has_transform) { out->position = transform.apply_point(out->position); out->normal = transform.apply_normal(out->normal); }
        if (has_animation) { out->position = anim.apply_point(out->position); out->normal = anim.apply_normal(out->normal); }
        return true;
    }
};

struct Bvh {  // :352-482
    std::vector<SceneItem> items;
    std::vector<BvhNode> nodes;
    size_t initial_index = 0;
    size_t depth = 0;

    void build() {  // :360-380
        std::vector<std::pair<size_t, Aabb>> h(items.size());
        for (size_t i = 0; i < items.size(); ++i) h[i] = {USIZE_MAX - i, items[i].bounding_box()};
        nodes.clear();
        size_t d = 0;
        Aabb box;
        initial_index = build_rec(h.data(), h.size(), 0, 0, &box, &d);
        depth = d;
    }
    size_t build_rec(std::pair<size_t, Aabb>* h, size_t n, size_t axis_id, size_t depth_in, Aabb* out_box,
                     size_t* out_depth) {  // :382-427
        if (n == 1) {
            *out_box = h[0].second;
            *out_depth = depth_in;
            return h[0].first;
        }
        std::stable_sort(h, h + n, [axis_id](const std::pair<size_t, Aabb>& a, const std::pair<size_t, Aabb>& b) {
            return a.second.min.e[axis_id] < b.second.min.e[axis_id];
        });
        if (n == 2) {
            size_t id = nodes.size();
            BvhNode nd;
            nd.aabb = aabb_surrounding(h[0].second, h[1].second);
            nd.left = h[0].first;
            nd.right = h[1].first;
            nd.axis_id = axis_id;
            nodes.push_back(nd);
            *out_box = nd.aabb;
            *out_depth = depth_in;
            return id;
        }
        size_t mid = n / 2;
        size_t id = nodes.size();
        nodes.push_back(BvhNode());
        Aabb lb, rb;
        size_t ld, rd;
        size_t l = build_rec(h, mid, (axis_id + mid) % 3, depth_in + 1, &lb, &ld);
        size_t r = build_rec(h + mid, n - mid, (axis_id + (n - mid)) % 3, depth_in + 1, &rb, &rd);
        BvhNode nd;
        nd.aabb = aabb_surrounding(lb, rb);
        nd.left = l;
        nd.right = r;
        nd.axis_id = axis_id;
        nodes[id] = nd;
        *out_box = nd.aabb;
        *out_depth = std::max(ld, rd);
        return id;
    }

    bool hit_recursive(size_t node, const Ray& ray, Range& t_range, TRng& rng, HitInteraction* out, Counters* cnt) const {  // :429-451
        if (node > items.size()) {
            size_t item = USIZE_MAX - node;
            if (cnt) cnt->prim_tests++;
            GeoHit g;
            if (items[item].hit(ray, t_range, rng, (uint32_t)item, &g)) {  // leaf: NO AABB test (Q13)
                out->position = g.position; out->normal = g.normal; out->uv = g.uv; out->t = g.t;
                out->front_face = g.front_face; out->material = items[item].material; out->item = (uint32_t)item;
                return true;
            }
            return false;
        }
        if (cnt) cnt->node_tests++;
        if (aabb_hit_cond(nodes[node].aabb, ray, t_range)) {
            size_t order[2];
            if (ray.direction.e[nodes[node].axis_id] > 0.0f) { order[0] = nodes[node].left; order[1] = nodes[node].right; }
            else { order[0] = nodes[node].right; order[1] = nodes[node].left; }
            bool any = false;  // hit_index_list :453-462
            for (int k = 0; k < 2; ++k) {
                HitInteraction hi;
                if (hit_recursive(order[k], ray, t_range, rng, &hi, cnt)) {
                    t_range.end = hi.t;
                    *out = hi;
                    any = true;
                }
            }
            return any;
        }
        return false;
    }
    bool hit(const Ray& ray, const Range& t_range_in, TRng& rng, HitInteraction* out, Counters* cnt) const {  // :464-473
        Range t_range = t_range_in;
        if (cnt) cnt->rays++;
        return hit_recursive(initial_index, ray, t_range, rng, out, cnt);
    }
};

// ---------------------------------------------------------------- src/lib/camera.rs
struct Camera {  // :154-164
    Point3 position;
    Dir3 upper_left_corner, unit_right, unit_up, scaled_right, scaled_up;
    float lens_radius = 0.0f;
    float t0 = 0.0f, t1 = 0.0f;
    Vec2f shutter_pace;
    float aspect_ratio() const { return length(scaled_up) / length(scaled_right); }  // :171-173

    Ray ray(TRng& rng, Vec2f point) const {  // :175-200
        Dir3 offset(0, 0, 0);
        if (lens_radius > 0.0f) {
            float rd[2];
            rng.unit_disc(rd);
            offset = lens_radius * (rd[0] * unit_right + rd[1] * unit_up);
        }
        float start_time = (t0 == t1) ? t0 : rng.gen_range(t0, t1);
        float time = start_time + (shutter_pace.x * point.x + shutter_pace.y * point.y);
        Ray r;
        r.origin = position + offset;
        r.direction = unit(upper_left_corner + point.x * scaled_right - point.y * scaled_up - offset);
        r.time = time;
        return r;
    }
};
struct CameraBuilder {  // :9-152
    float viewport_width = 0, viewport_height = 0;
    Point3 position;
    Dir3 up = Dir3(0, 1, 0), forward = Dir3(0, 0, -1);
    float aperture = 0.0f, focus_distance = 1.0f;
    float t0 = 0, t1 = 0;
    CameraBuilder& vertical_fov(float fov_deg, float aspect_ratio) {  // :72-79 (Q10)
        float h = rt_tanf((fov_deg * (PI_F / 180.0f)) * 0.5f);
        viewport_width = 2.0f * h;
        viewport_height = 2.0f * h * aspect_ratio;
        return *this;
    }
    CameraBuilder& set_position(Point3 p) { position = p; return *this; }
    CameraBuilder& look_at(Dir3 up_, Point3 pos) { up = up_; forward = pos - position; return *this; }  // :95-99
    CameraBuilder& look_at_focus(Dir3 up_, Point3 pos) {                                                // :100-105
        up = up_; forward = pos - position; focus_distance = length(forward); return *this;
    }
    CameraBuilder& set_aperture(float a) { aperture = a; return *this; }
    CameraBuilder& set_focus_distance(float d) { focus_distance = d; return *this; }
    CameraBuilder& motion_blur(float a, float b) { t0 = a; t1 = b; return *this; }
    Camera build() const {  // :131-152
        Camera c;
        c.unit_right = unit(cross(forward, up));
        c.unit_up = unit(cross(c.unit_right, forward));
        Dir3 sforward = with_length(forward, focus_distance);
        c.upper_left_corner =
            focus_distance * (c.unit_right * (viewport_width * -0.5f) + c.unit_up * (viewport_height * 0.5f)) + sforward;
        c.position = position;
        c.scaled_right = c.unit_right * (focus_distance * viewport_width);
        c.scaled_up = c.unit_up * (focus_distance * viewport_height);
        c.lens_radius = aperture / 2.0f;
        c.t0 = t0; c.t1 = t1;
        return c;
    }
};

// ---------------------------------------------------------------- src/lib/background_color.rs
struct BackgroundColor {
    bool sky = true;
    Color color;
    Color sample(const Ray& ray) const {  // :9-19
        if (sky) {
            float t = 0.5f * (dot(DIR_UP, ray.direction) + 1.0f);
            return lerp(Color(1.0f, 1.0f, 1.0f), Color(0.5f, 0.7f, 1.0f), t);
        }
        return color;
    }
};

// ---------------------------------------------------------------- src/lib/rendering.rs
struct World {  // :12-17
    Camera camera;
    Bvh scene;
    BackgroundColor background;
    bool has_light = false;  // scattering_distribution_provider: Option<Rect>
    RectGeometry light;
};

static inline Color ray_color(const Ray& ray, const World& world, TRng& rng, int max_depth, Counters* cnt) {  // :19-71
    int depth = max_depth;
    Color accum_att(1, 1, 1);
    Color accum_emitted(0, 0, 0);
    Ray cur_ray = ray;
    uint32_t bounce = 0;
    for (;;) {
        rng.begin_block(1u + bounce);  // REPLAY bookkeeping only (no-op in FAITHFUL mode)
        ++bounce;
        HitInteraction hi;
        Range tr{0.001f, INF_F};
        if (world.scene.hit(cur_ray, tr, rng, &hi, cnt)) {
            if (depth <= 1) return Color(0, 0, 0);  // Q7
            Color att;
            ScatterDist dist;
            if (material_scatter(*hi.material, cur_ray, hi, rng, &att, &dist)) {
                Ray scattered;
                float probability;
                if (dist.mirror) {
                    scattered = Ray{hi.position, dist.generate(rng), cur_ray.time};
                    probability = 1.0f;
                } else {
                    float scattered_pdf;
                    Dir3 dir;
                    if (world.has_light) {  // sample_final_scattering_distribution :73-92
                        const float mix = 0.5f;
                        if (rng.gen_bool_half()) dir = rect_generate(world.light, hi.position, rng);
                        else dir = dist.generate(rng);
                        scattered_pdf = mix * rect_value(world.light, hi.position, dir) + (1.0f - mix) * dist.value(dir);
                    } else {
                        dir = dist.generate(rng);
                        scattered_pdf = dist.value(dir);
                    }
                    scattered = Ray{hi.position, dir, cur_ray.time};
                    float scattering_pdf = material_scattering_pdf(*hi.material, scattered, hi);
                    probability = scattering_pdf / scattered_pdf;  // Q9: 0/0 = NaN
                }
                Color emitted = material_emit(*hi.material, hi);
                accum_emitted = accum_emitted + convolution(accum_att, emitted);
                accum_att = convolution(accum_att, att) * probability;
                cur_ray = scattered;
                depth -= 1;
                continue;
            }
            Color emitted = material_emit(*hi.material, hi);
            return accum_emitted + convolution(accum_att, emitted);
        }
        Color emitted = world.background.sample(ray);  // Q6: PRIMARY ray
        return accum_emitted + convolution(accum_att, emitted);
    }
}

// split_work_tasks :222-237
static inline std::vector<std::pair<size_t, size_t>> split_work_tasks(size_t spp, size_t thread_count) {
    size_t whole = spp / thread_count, rem = spp % thread_count;
    std::vector<std::pair<size_t, size_t>> tasks;
    for (size_t id = 0; id < thread_count; ++id) {
        size_t real = whole + (id < rem ? 1 : 0);
        if (real == 0) break;
        tasks.push_back({id, real});
    }
    return tasks;
}

}  // namespace orc
