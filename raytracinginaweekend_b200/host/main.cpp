// rtiaw_app — the reference CLI (src/app/main.rs:10-64) on top of the B200 render path.
//
//   rtiaw_app <out.png|out.ppm> <image_width> <samples_per_pixel> <max_depth> <demo:name | config:cN>
//
// Same positional arguments, same height rule (`(width as f32 * camera.aspect_ratio()) as i32`, main.rs:38-41), same
// gamma-2 quantisation (Color::to_rgb8_gamma2, color.rs:43-48).  `thread_count` = available_parallelism in the reference
// (main.rs:19); here it is the number of sample lanes in flight per pixel (env RTIAW_LANES, default 16).
// The reference seeds its render RNG from OS entropy (rendering.rs:160); so does this CLI unless RTIAW_SEED is set (tests).
// Assets are looked up under $RTIAW_ASSETS (default ./assets); worlds with an image texture read <assets>/earthmap.ppm
// (P6; `python -m raytracinginaweekend_b200.tools_earth_ppm` writes it from earthmap.jpg — JPEG decode is caller-side, the
// reference delegates it to the `image` crate, world_builder.rs:19-31).  Without that file those worlds fail with a clear message.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>

#include "rtiaw_app.hpp"

using namespace rtiaw;

static uint32_t crc_table[256];
static void crc_init() {
    for (uint32_t n = 0; n < 256; ++n) {
        uint32_t c = n;
        for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
        crc_table[n] = c;
    }
}
static uint32_t crc32(uint32_t c, const uint8_t* p, size_t n) {
    c = ~c;
    for (size_t i = 0; i < n; ++i) c = crc_table[(c ^ p[i]) & 0xFF] ^ (c >> 8);
    return ~c;
}
static void be32(std::vector<uint8_t>& v, uint32_t x) { for (int s = 24; s >= 0; s -= 8) v.push_back((uint8_t)(x >> s)); }
static void chunk(std::vector<uint8_t>& out, const char* type, const std::vector<uint8_t>& data) {
    be32(out, (uint32_t)data.size());
    std::vector<uint8_t> td(type, type + 4);
    td.insert(td.end(), data.begin(), data.end());
    out.insert(out.end(), td.begin(), td.end());
    be32(out, crc32(0, td.data(), td.size()));
}
// Minimal PNG writer: 8-bit RGB, zlib stream of stored (uncompressed) deflate blocks.
static bool write_png(const std::string& path, int w, int h, const std::vector<uint8_t>& rgb) {
    crc_init();
    std::vector<uint8_t> raw;
    raw.reserve((size_t)h * (3 * w + 1));
    for (int y = 0; y < h; ++y) {
        raw.push_back(0);
        raw.insert(raw.end(), rgb.begin() + (size_t)y * 3 * w, rgb.begin() + (size_t)(y + 1) * 3 * w);
    }
    std::vector<uint8_t> z = {0x78, 0x01};
    uint32_t a = 1, b = 0;
    for (uint8_t c : raw) { a = (a + c) % 65521; b = (b + a) % 65521; }
    for (size_t off = 0; off < raw.size(); off += 65535) {
        const size_t n = std::min<size_t>(65535, raw.size() - off);
        z.push_back(off + n >= raw.size() ? 1 : 0);
        z.push_back((uint8_t)(n & 0xFF)); z.push_back((uint8_t)(n >> 8));
        z.push_back((uint8_t)(~n & 0xFF)); z.push_back((uint8_t)((~n >> 8) & 0xFF));
        z.insert(z.end(), raw.begin() + off, raw.begin() + off + n);
    }
    be32(z, (b << 16) | a);
    std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A}, ihdr;
    be32(ihdr, (uint32_t)w); be32(ihdr, (uint32_t)h);
    ihdr.insert(ihdr.end(), {8, 2, 0, 0, 0});
    chunk(out, "IHDR", ihdr);
    chunk(out, "IDAT", z);
    chunk(out, "IEND", {});
    std::ofstream f(path, std::ios::binary);
    f.write((const char*)out.data(), (std::streamsize)out.size());
    return (bool)f;
}
static bool read_ppm(const std::string& path, uint32_t& w, uint32_t& h, std::vector<uint8_t>& rgb) {
    std::ifstream f(path, std::ios::binary);
    std::string magic;
    int maxv = 0;
    if (!(f >> magic >> w >> h >> maxv) || magic != "P6" || maxv != 255) return false;
    f.get();
    rgb.resize((size_t)w * h * 3);
    f.read((char*)rgb.data(), (std::streamsize)rgb.size());
    return (bool)f;
}

int main(int argc, char** argv) {
    if (argc < 6) {
        std::fprintf(stderr, "usage: %s <out.png|out.ppm> <image_width> <samples_per_pixel> <max_depth> <demo:name|config:cN>\n", argv[0]);
        return 2;
    }
    try {
        const std::string out_path = argv[1];
        const int image_width = std::atoi(argv[2]);
        const size_t spp = (size_t)std::atoll(argv[3]);
        const int max_depth = std::atoi(argv[4]);
        const std::string world_name = argv[5];
        const char* ad = std::getenv("RTIAW_ASSETS");
        const std::string assets_dir = ad ? ad : "assets";
        const char* ln = std::getenv("RTIAW_LANES");
        const size_t thread_count = ln ? (size_t)std::atoll(ln) : 16;
        std::fprintf(stderr, "Using %zu sample lanes.\n", thread_count);
        Assets a;
        a.suzanne_obj = assets_dir + "/suzanne.obj";
        a.cube_obj = assets_dir + "/cube.obj";
        std::vector<uint8_t> earth;
        if (read_ppm(assets_dir + "/earthmap.ppm", a.earth_w, a.earth_h, earth)) a.earth_rgb8 = earth.data();
        const bool needs_earth = world_name == "demo:earth_mapped" || world_name == "demo:final_scene2" || world_name == "config:c3";
        if (needs_earth && !a.earth_rgb8)
            throw std::runtime_error("world " + world_name + " needs the earth map: " + assets_dir + "/earthmap.ppm is missing "
                                     "(write it with: python -m raytracinginaweekend_b200.tools_earth_ppm)");
        NamedWorld nw = create_world(world_name, a);
        const Size2i size{image_width, (int32_t)((float)image_width * nw.world->camera.aspect_ratio())};
        std::fprintf(stderr, "Start rendering...\n");
        RtStats st{};
        const auto t0 = std::chrono::steady_clock::now();
        const char* sd = std::getenv("RTIAW_SEED");
        const uint64_t seed = sd ? (uint64_t)std::strtoull(sd, nullptr, 10) : (uint64_t)std::chrono::system_clock::now().time_since_epoch().count();
        std::vector<Color> px = render(size, thread_count, spp, max_depth, *nw.world, RenderMode::Default, seed, &st);
        const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        std::fprintf(stderr, "Rendering done in %f seconds (%.1f Mrays/s on the device)\n", secs, st.ray_segments / (st.render_ms * 1e3));
        std::fprintf(stderr, "Saving image...\n");
        std::vector<uint8_t> bytes(px.size() * 3);
        for (size_t i = 0; i < px.size(); ++i) to_rgb8_gamma2(px[i], &bytes[3 * i]);
        if (out_path.size() > 4 && out_path.substr(out_path.size() - 4) == ".ppm") {
            std::ofstream f(out_path, std::ios::binary);
            f << "P6\n" << size.width << " " << size.height << "\n255\n";
            f.write((const char*)bytes.data(), (std::streamsize)bytes.size());
        } else if (!write_png(out_path, size.width, size.height, bytes)) {
            throw std::runtime_error("cannot write " + out_path);
        }
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
}
