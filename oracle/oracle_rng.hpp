// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.
// PARITY UNPINNED: the reference has no tests / golden vectors and cannot be compiled here
// (no rustc); the third-party RNG crates are restated from their published algorithms.
//
// Random source of the CPU oracle.  Two modes:
//   * FAITHFUL  — sequential Xoroshiro128++ with the `rand 0.8.5` / `rand_distr 0.4.3`
//                 sampling algorithms the reference calls (rejection samplers).
//                 Call sites: src/lib/rendering.rs:80,132-138,160,170,174; src/lib/camera.rs:178,188;
//                 src/lib/material.rs:21,67,97; src/lib/hittable/rect_geometry.rs:63-64;
//                 src/lib/common.rs:1 (TRng = rand_xoshiro::Xoroshiro128PlusPlus).
//   * REPLAY    — counter-based Philox4x32-10 keyed (seed, pixel, sample, draw-block) with
//                 inversion samplers.  Distribution-equivalent to FAITHFUL (render RNG is
//                 entropy-seeded in the reference, src/lib/rendering.rs:160, so only the
//                 distribution is observable) and laid out exactly like the GPU path so
//                 GPU and oracle walk the same light paths.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <utility>

namespace orc {

static inline uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }

static inline float bits_to_f32(uint32_t b) {
    float f;
    std::memcpy(&f, &b, 4);
    return f;
}
static inline uint32_t f32_to_bits(float f) {
    uint32_t b;
    std::memcpy(&b, &f, 4);
    return b;
}

// ---- Philox4x32-10 (Salmon et al., SC'11) -------------------------------------------------
static inline void philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2,
                                 uint32_t c3, uint32_t out[4]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)M0 * c0;
        uint64_t p1 = (uint64_t)M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// Deterministic sin/cos of 2*pi*u, u in [0,1): plain IEEE + - * only, so the same code gives
// the same bits on the CPU (no contraction) and on the GPU (--fmad=false).
static inline void sincos_2pi(float u, float* s_out, float* c_out) {
    float x = u * 4.0f;
    int q = (int)x;  // 0..3
    float f = x - (float)q;
    float a = f * 1.57079632679489661923f;
    float a2 = a * a;
    float s = a * (1.0f + a2 * (-1.6666667163e-1f + a2 * (8.3333337680e-3f + a2 * (-1.9841270114e-4f +
              a2 * (2.7557314297e-6f + a2 * (-2.5052107943e-8f))))));
    float c = 1.0f + a2 * (-0.5f + a2 * (4.1666667908e-2f + a2 * (-1.3888889225e-3f + a2 * (2.4801587642e-5f +
              a2 * (-2.7557319224e-7f + a2 * (2.0876755880e-9f))))));
    switch (q & 3) {
        case 0: *c_out = c;  *s_out = s;  break;
        case 1: *c_out = -s; *s_out = c;  break;
        case 2: *c_out = -c; *s_out = -s; break;
        default: *c_out = s; *s_out = -c; break;
    }
}

// Deterministic cube root on [0,1): bit-hack seed + 4 Newton steps, IEEE ops only.
static inline float cbrt_unit(float u) {
    if (!(u > 0.0f)) return 0.0f;
    float y = bits_to_f32(f32_to_bits(u) / 3u + 709921077u);
    for (int i = 0; i < 4; ++i) y = (2.0f * y + u / (y * y)) * 0.333333343267f;
    return y;
}

// Deterministic natural log on [0,1): exponent split + atanh series, IEEE ops only (same bits on CPU and GPU).
static inline float ln_unit(float u) {
    if (!(u > 0.0f)) return -std::numeric_limits<float>::infinity();
    uint32_t b = f32_to_bits(u);
    int e = (int)(b >> 23) - 127;
    float m = bits_to_f32((b & 0x007FFFFFu) | 0x3F800000u);   // [1, 2)
    if (m > 1.41421356f) { m = m * 0.5f; e += 1; }           // [0.707, 1.414)
    float s = (m - 1.0f) / (m + 1.0f);                        // ln m = 2 atanh(s), |s| < 0.172
    float s2 = s * s;
    float p = 2.0f * s * (1.0f + s2 * (0.333333343f + s2 * (0.2f + s2 * (0.142857149f + s2 * 0.111111112f))));
    return (float)e * 0.693147182f + p;
}

// rand 0.8.5 `UniformFloat<f32>::new(low, high)` (used for the pixel jitter via src/lib/vec2.rs:43-83
// and inside rand_distr's rejection samplers).
struct UniformF32 {
    float low, scale;
    UniformF32() : low(0), scale(1) {}
    UniformF32(float lo, float hi) {
        low = lo;
        scale = hi - lo;
        const float max_rand = 1.0f - 1.1920929e-7f;  // 1 - f32::EPSILON
        while (scale * max_rand + low >= hi) scale = bits_to_f32(f32_to_bits(scale) - 1);
    }
};

struct TRng {
    // FAITHFUL state (rand_xoshiro 0.6.0 Xoroshiro128PlusPlus)
    uint64_t s0 = 1, s1 = 2;
    // REPLAY state
    bool replay = false;
    uint32_t key0 = 0, key1 = 0, pixel = 0, sample = 0;
    uint32_t buf[4] = {0, 0, 0, 0};
    uint32_t block_id = 0;
    int idx = 4;

    static TRng from_seed_bytes(const uint8_t seed[16]) {
        TRng r;
        std::memcpy(&r.s0, seed, 8);      // little-endian u64 x2
        std::memcpy(&r.s1, seed + 8, 8);
        return r;
    }
    static TRng from_u64(uint64_t seed) {  // SplitMix64 expansion (oracle's own; the reference seeds from entropy)
        TRng r;
        auto next = [&seed]() {
            seed += 0x9E3779B97F4A7C15ull;
            uint64_t z = seed;
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
            z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
            return z ^ (z >> 31);
        };
        r.s0 = next();
        r.s1 = next();
        return r;
    }
    static TRng make_replay(uint64_t seed) {
        TRng r;
        r.replay = true;
        r.key0 = (uint32_t)seed;
        r.key1 = (uint32_t)(seed >> 32);
        return r;
    }

    // REPLAY only: select the (pixel, sample) stream and the draw block (0 = camera, 1+b = bounce b).
    void set_path(uint32_t pix, uint32_t smp) { pixel = pix; sample = smp; }
    void begin_block(uint32_t block) {
        if (!replay) return;
        block_id = block;
        philox4x32_10(key0, key1, pixel, sample, block, 0u, buf);
        idx = 0;
    }

    uint64_t next_u64() {
        uint64_t r = rotl64(s0 + s1, 17) + s0;
        uint64_t t = s1 ^ s0;
        s0 = rotl64(s0, 49) ^ t ^ (t << 21);
        s1 = rotl64(t, 28);
        return r;
    }
    uint32_t next_u32() { return (uint32_t)next_u64(); }

    // uniform in [0,1) with 24 bits; in REPLAY mode this is the only primitive.
    float u01() {
        if (replay) {
            if (idx > 0 && (idx & 3) == 0) philox4x32_10(key0, key1, pixel, sample, block_id, (uint32_t)(idx >> 2), buf);  // draws 4.. of a block
            uint32_t v = buf[idx & 3];
            ++idx;
            return (float)(v >> 8) * 5.9604644775390625e-8f;
        }
        return (float)(next_u32() >> 8) * 5.9604644775390625e-8f;
    }
    // rand `Standard` f32: (u32 >> 8) * 2^-24
    float gen_f32() { return u01(); }
    // rand `gen_bool(0.5)`: next_u64 < 2^63   (src/lib/rendering.rs:80)
    bool gen_bool_half() {
        if (replay) return u01() < 0.5f;
        return next_u64() < (1ull << 63);
    }
    // rand `Uniform::sample`
    float uniform(const UniformF32& d) {
        if (replay) return u01() * d.scale + d.low;
        float v12 = bits_to_f32((next_u32() >> 9) | 0x3F800000u);
        return (v12 - 1.0f) * d.scale + d.low;
    }
    // rand `gen_range(lo..hi)` (sample_single)
    float gen_range(float lo, float hi) {
        if (replay) return lo + u01() * (hi - lo);
        float scale = hi - lo;
        for (;;) {
            float v12 = bits_to_f32((next_u32() >> 9) | 0x3F800000u);
            float res = v12 * scale + (lo - scale);
            if (res < hi) return res;
            scale = bits_to_f32(f32_to_bits(scale) - 1);
        }
    }
    // rand 0.8.5 UniformInt<u32>::sample_single(0, n) (widening-multiply rejection with the leading-zeros zone); used by
    // SliceRandom::shuffle -> gen_index (src/lib/perlin.rs:95).  Scene construction only (FAITHFUL stream).
    uint32_t gen_index(uint32_t ubound) {
        const uint32_t range = ubound;  // high - low with low = 0, exclusive upper bound
        if (range == 0) return next_u32();
        const uint32_t zone = (range << __builtin_clz(range)) - 1u;
        for (;;) {
            const uint64_t m = (uint64_t)next_u32() * range;
            if ((uint32_t)m <= zone) return (uint32_t)(m >> 32);
        }
    }
    template <class T>
    void shuffle(T* v, size_t n) {  // rand::seq::SliceRandom::shuffle
        for (size_t i = n - 1; i >= 1; --i) std::swap(v[i], v[gen_index((uint32_t)(i + 1))]);
    }
    // -ln(U) * (1/density) free-path draw of VolumeGeometry::hit (src/lib/hittable.rs:328): `rng.gen::<f32>().ln()`.
    // REPLAY: the draw is keyed by the element index (a ray tests a leaf at most once), so it does not depend on visit order;
    // ln is evaluated with ln_unit (IEEE ops only) so CPU and GPU agree bit for bit.  FAITHFUL: next sequential draw, libm logf.
    float volume_ln_u(uint32_t item) {
        if (replay) {
            uint32_t o[4];
            philox4x32_10(key0, key1, pixel, sample, block_id, 0x40000000u | item, o);
            return ln_unit((float)(o[0] >> 8) * 5.9604644775390625e-8f);
        }
        return std::log(gen_f32());
    }
    // rand_distr 0.4.3 UnitSphere (Marsaglia 1972)  — src/lib/material.rs:21
    void unit_sphere(float out[3]) {
        if (replay) {
            float ua = u01(), ub = u01();
            float z = 1.0f - 2.0f * ua;
            float r2 = 1.0f - z * z;
            float r = std::sqrt(r2 > 0.0f ? r2 : 0.0f);
            float s, c;
            sincos_2pi(ub, &s, &c);
            out[0] = r * c; out[1] = r * s; out[2] = z;
            return;
        }
        UniformF32 d(-1.0f, 1.0f);
        for (;;) {
            float x1 = uniform(d), x2 = uniform(d);
            float sum = x1 * x1 + x2 * x2;
            if (sum >= 1.0f) continue;
            float factor = 2.0f * std::sqrt(1.0f - sum);
            out[0] = x1 * factor; out[1] = x2 * factor; out[2] = 1.0f - 2.0f * sum;
            return;
        }
    }
    // rand_distr 0.4.3 UnitBall — src/lib/material.rs:67
    void unit_ball(float out[3]) {
        if (replay) {
            float dir[3];
            unit_sphere(dir);
            float rad = cbrt_unit(u01());
            out[0] = dir[0] * rad; out[1] = dir[1] * rad; out[2] = dir[2] * rad;
            return;
        }
        UniformF32 d(-1.0f, 1.0f);
        for (;;) {
            float x1 = uniform(d), x2 = uniform(d), x3 = uniform(d);
            if (x1 * x1 + x2 * x2 + x3 * x3 <= 1.0f) { out[0] = x1; out[1] = x2; out[2] = x3; return; }
        }
    }
    // rand_distr 0.4.3 UnitDisc — src/lib/camera.rs:178
    void unit_disc(float out[2]) {
        if (replay) {
            float ua = u01(), ub = u01();
            float r = std::sqrt(ua);
            float s, c;
            sincos_2pi(ub, &s, &c);
            out[0] = r * c; out[1] = r * s;
            return;
        }
        UniformF32 d(-1.0f, 1.0f);
        for (;;) {
            float x1 = uniform(d), x2 = uniform(d);
            if (x1 * x1 + x2 * x2 <= 1.0f) { out[0] = x1; out[1] = x2; return; }
        }
    }
};

}  // namespace orc
