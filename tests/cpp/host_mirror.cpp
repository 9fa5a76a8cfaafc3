// Drives the C++ host mirror (include/sar.hpp) the way the reference's only in-repo caller does
// (src/bin/main.rs:482-518): single-thread path Runtime::new → render → colorize → reset, and
// the parallel path ParallelRenderer::new → render_parallel → shutdown.  Prints FNV-1a hashes of
// the images so the Python test can compare them with the same calls made through api.py.
#include <cstdio>
#include <cstring>

#include "sar.hpp"

static unsigned long long fnv_bytes(const unsigned char *p, size_t n)
{
    unsigned long long h = 1469598103934665603ull;
    for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}
static unsigned long long fnv(const std::vector<uint16_t> &v) { return fnv_bytes(reinterpret_cast<const unsigned char *>(v.data()), v.size() * 2); }

int main(int argc, char **argv)
{
    using namespace sar;
    Config cfg = Config::poisson_saturne();
    Config sol = Config::solar_sail();
    // presets round-trip through the POD
    const sar_config a = cfg.to_pod(), b = Config::from_pod(a).to_pod();
    if (std::memcmp(&a, &b, sizeof a) != 0) { std::puts("POD_ROUNDTRIP_FAILED"); return 2; }
    if (!std::holds_alternative<color_transforms::AdjustedVelocity>(sol.color_transform) || sol.view.scale != 1.7) { std::puts("PRESET_FAILED"); return 2; }
    std::puts("PRESETS_OK");

    if (argc > 1 && std::strcmp(argv[1], "--no-gpu") == 0) {
        try {
            Runtime rt(cfg);
            std::puts("UNEXPECTED_RUNTIME");
            return 3;
        } catch (const Error &e) {
            std::printf("NO_GPU_ERROR code=%d\n", e.code);
            return e.code == SAR_ERR_CUDA ? 0 : 4;
        }
    }

    // single-thread path, main.rs:482-491
    cfg.iterations = 200000; cfg.width = 320; cfg.height = 200;
    Runtime runtime(cfg, 0, /*seed=*/42);
    for (int frame = 0; frame < 2; ++frame) {
        cfg.angle = 0.3 * frame;
        render(cfg, runtime);
        FinalImage image = colorize(cfg, runtime);
        std::printf("SINGLE frame=%d hash=%llu\n", frame, fnv(image.raw));
        runtime.reset();
    }
    // merge panics on a dimension mismatch (lib.rs:709-710)
    Config other = cfg; other.width = 100;
    Runtime small(other, 0, 1);
    try { runtime.merge(small); std::puts("MERGE_DID_NOT_FAIL"); return 5; }
    catch (const Error &e) { std::printf("MERGE_MISMATCH code=%d\n", e.code); }

    // parallel path, main.rs:492-518
    sol.iterations = 3000000; sol.width = 180; sol.height = 200; sol.angle = 3.839724354387525;
    ParallelRenderer renderer({}, 256);
    FinalImage img = render_parallel(renderer, sol, 12, /*seed=*/7);
    std::printf("PARALLEL threads=%llu hash=%llu px00=%u,%u,%u\n", (unsigned long long)renderer.num_threads(), fnv(img.raw),
                img.pixel(0, 0)[0], img.pixel(0, 0)[1], img.pixel(0, 0)[2]);
    renderer.shutdown();

    // write_image_matches (main.rs:40-100) and the auto-framing first pass (lib.rs:326-334)
    cfg.angle = 0.0; cfg.transparent = false;
    render(cfg, runtime);
    (void)colorize(cfg, runtime);
    const std::vector<uint8_t> pam = encode_image(runtime, cfg.width, cfg.height, pixel_format(cfg.transparent, false), Container::Pam);
    const std::vector<uint8_t> bmp = encode_image(runtime, cfg.width, cfg.height, pixel_format(cfg.transparent, true), Container::Bmp);
    const std::vector<uint8_t> png = encode_png(runtime, cfg.width, cfg.height, pixel_format(cfg.transparent, false));
    std::printf("ENCODED pam=%llu bmp=%llu png=%llu\n", fnv_bytes(pam.data(), pam.size()), fnv_bytes(bmp.data(), bmp.size()),
                fnv_bytes(png.data(), png.size()));
    try { (void)encode_image(runtime, cfg.width, cfg.height, PixelFormat::Rgb16, Container::Bmp); std::puts("BMP16_DID_NOT_FAIL"); return 6; }
    catch (const Error &e) { std::printf("BMP16 code=%d\n", e.code); }
    // the binary's frame loop (main.rs:496-512) through the mirror: raw RGBA16 frames and complete compressed PNG files
    {
        Config seq = Config::solar_sail();
        seq.iterations = 600000; seq.width = 160; seq.height = 120; seq.transparent = false;
        ParallelRenderer r2({}, 512);
        const std::vector<double> angles = angle_iter(0.0, 50.0, 10.0);
        unsigned long long h16 = 1469598103934665603ull, hpng = 1469598103934665603ull;
        size_t n16 = 0, npng = 0, png_bytes = 0;
        render_sequence(r2, seq, angles, 2, /*seed=*/5, [&](uint32_t f, const uint16_t *px) {
            (void)f; ++n16;
            const uint8_t *b = reinterpret_cast<const uint8_t *>(px);
            for (size_t i = 0; i < size_t(seq.width) * seq.height * 8; ++i) h16 = (h16 ^ b[i]) * 1099511628211ull;
        });
        render_sequence_encoded(r2, seq, angles, 2, /*seed=*/5, PixelFormat::Rgb16, Container::PngDeflate,
                                [&](uint32_t f, const uint8_t *b, size_t n) {
                                    (void)f; ++npng; png_bytes += n;
                                    for (size_t i = 0; i < n; ++i) hpng = (hpng ^ b[i]) * 1099511628211ull;
                                });
        std::printf("SEQUENCE frames=%zu hash=%llu pngframes=%zu pngbytes=%zu pnghash=%llu first=%.17g\n", n16, h16, npng, png_bytes, hpng, angles[0]);
    }
    const AutoFrame af = autoframe(Config::poisson_saturne(), 1024, 2000, /*seed=*/3);
    std::printf("AUTOFRAME xmin=%.17g ymax=%.17g diverged=%llu\n", af.box[0], af.box[3], (unsigned long long)af.diverged);
    return 0;
}
