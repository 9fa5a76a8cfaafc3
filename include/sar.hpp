// sar.hpp — C++17 host mirror of the reference's public API for the render path, header-only,
// over the C ABI of sar.h.  The reference is a Rust library; this is the same surface for
// compiled callers where no Rust toolchain exists (names, argument meaning and failure points
// follow src/lib.rs; where the reference panics, sar::Error is thrown):
//
//   sar::Vec3, EulerAxisRotation, View, RenderKind, BrighnessConstants, Palette, Colors   lib.rs:115-175, 232-492
//   sar::attractors::PolynomialSprott2Degree                                               lib.rs:575-580
//   sar::color_transforms::{PoissonSaturne, AdjustedVelocity}                              lib.rs:503-559
//   sar::Config::{poisson_saturne, solar_sail}                                             lib.rs:310-387
//   sar::Runtime::{Runtime(config), reset, merge}                                          lib.rs:660, 682, 708
//   sar::render(config, runtime), sar::colorize(config, runtime) -> FinalImage             lib.rs:747, 841
//   sar::ParallelRenderer::{ParallelRenderer(), shutdown}, sar::render_parallel(...)       lib.rs:919, 1020, 1051
//   sar::PixelFormat, Container, encode_image(runtime, ...), encode_png(...), write_image(...)  src/bin/main.rs:40-100
//   sar::angle_iter, render_sequence(...), render_sequence_encoded(...)                    src/bin/main.rs:107-176, 496-512
//   sar::autoframe(config, ...) -> AutoFrame                                               lib.rs:326-334 (the author's TODO)
#pragma once
#include <array>
#include <cstdint>
#include <random>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <variant>
#include <vector>

#include "sar.h"

namespace sar {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error("sar error " + std::to_string(c) + ": " + m), code(c) {}
};
inline void check(int rc) { if (rc != SAR_OK) throw Error(rc, sar_last_error()); }

struct Vec3 { double x, y, z; };                                   // lib.rs:115-119
struct EulerAxisRotation { Vec3 axis; double rotation; };          // lib.rs:170-175 (axis used as given, lib.rs:181-183)
struct View { Vec3 center_camera; EulerAxisRotation rotation; double scale; };   // lib.rs:253-261
enum class RenderKind { Gas = SAR_RENDER_GAS, Depth = SAR_RENDER_DEPTH };          // lib.rs:234-239
struct BrighnessConstants { double offset = -0.15, factor = 5. / 3.; };           // lib.rs:390-404 (sic)

struct Palette {                                                   // lib.rs:408-473
    std::vector<std::array<double, 3>> list;                       // without the duplicated sentinel of lib.rs:418
    explicit Palette(std::vector<std::array<double, 3>> l) : list(std::move(l)) {
        if (list.empty()) throw Error(SAR_ERR_INVALID, "Palette::new panics if list.is_empty() (lib.rs:415)");
        if (list.size() > SAR_MAX_PALETTE) throw Error(SAR_ERR_INVALID, "too many palette entries for the C ABI");
    }
    template <size_t LEN>
    static Palette from_rgb(const std::array<double, LEN> &r, const std::array<double, LEN> &g, const std::array<double, LEN> &b) {
        std::vector<std::array<double, 3>> l;
        for (size_t i = 0; i < LEN; ++i) l.push_back({r[i], g[i], b[i]});
        return Palette(std::move(l));
    }
    size_t count() const { return list.size(); }
};
struct Colors {                                                    // lib.rs:475-492
    Palette palette = Palette::from_rgb<6>({1., 0.5, 1., 0.5, 0.5, 1.}, {1., 1., 0.5, 1., 0.5, 0.5}, {0.5, 0.5, 0.5, 1., 1., 1.});
    BrighnessConstants brighness;
};

namespace attractors {
struct PolynomialSprott2Degree { std::array<double, 10> x, y, z; };  // lib.rs:575-580
}
namespace color_transforms {
struct PoissonSaturne {};                                          // the fn item color_transforms::poisson_saturne, lib.rs:520
struct AdjustedVelocity { double offset, factor; };               // lib.rs:507-510
}
using ColorTransform = std::variant<color_transforms::PoissonSaturne, color_transforms::AdjustedVelocity>;

struct Config {                                                    // lib.rs:265-287, defaults lib.rs:289-307
    size_t iterations = 10'000'000;
    uint32_t width = 1920, height = 1080;
    RenderKind render = RenderKind::Gas;
    bool transparent = true;
    double angle = 0.0;                                            // radians (lib.rs:745)
    bool silent = true;
    attractors::PolynomialSprott2Degree attractor{};
    Colors colors;
    View view{};
    ColorTransform color_transform;

    static Config from_pod(const sar_config &c) {
        Config k;
        k.iterations = c.iterations; k.width = c.width; k.height = c.height;
        k.render = static_cast<RenderKind>(c.render_kind); k.transparent = c.transparent != 0; k.angle = c.angle; k.silent = c.silent != 0;
        for (int i = 0; i < 10; ++i) { k.attractor.x[i] = c.coef[0][i]; k.attractor.y[i] = c.coef[1][i]; k.attractor.z[i] = c.coef[2][i]; }
        k.view = View{{c.center_camera[0], c.center_camera[1], c.center_camera[2]}, {{c.axis[0], c.axis[1], c.axis[2]}, c.rotation}, c.scale};
        if (c.ct_kind == SAR_CT_ADJUSTED_VELOCITY) k.color_transform = color_transforms::AdjustedVelocity{c.ct_offset, c.ct_factor};
        else k.color_transform = color_transforms::PoissonSaturne{};
        std::vector<std::array<double, 3>> l;
        for (uint32_t i = 0; i < c.palette_len; ++i) l.push_back({c.palette_rgb[i][0], c.palette_rgb[i][1], c.palette_rgb[i][2]});
        k.colors.palette = Palette(std::move(l));
        k.colors.brighness = {c.bright_offset, c.bright_factor};
        return k;
    }
    static Config poisson_saturne() { sar_config c; check(sar_config_poisson_saturne(&c)); return from_pod(c); }   // lib.rs:310
    static Config solar_sail() { sar_config c; check(sar_config_solar_sail(&c)); return from_pod(c); }             // lib.rs:355

    sar_config to_pod() const {
        sar_config c{};
        c.iterations = iterations; c.width = width; c.height = height;
        c.render_kind = static_cast<uint32_t>(render); c.transparent = transparent; c.silent = silent; c.angle = angle;
        for (int i = 0; i < 10; ++i) { c.coef[0][i] = attractor.x[i]; c.coef[1][i] = attractor.y[i]; c.coef[2][i] = attractor.z[i]; }
        c.center_camera[0] = view.center_camera.x; c.center_camera[1] = view.center_camera.y; c.center_camera[2] = view.center_camera.z;
        c.axis[0] = view.rotation.axis.x; c.axis[1] = view.rotation.axis.y; c.axis[2] = view.rotation.axis.z;
        c.rotation = view.rotation.rotation; c.scale = view.scale;
        if (auto *av = std::get_if<color_transforms::AdjustedVelocity>(&color_transform)) {
            c.ct_kind = SAR_CT_ADJUSTED_VELOCITY; c.ct_offset = av->offset; c.ct_factor = av->factor;
        } else c.ct_kind = SAR_CT_POISSON_SATURNE;
        c.palette_len = static_cast<uint32_t>(colors.palette.count());
        for (uint32_t i = 0; i < c.palette_len; ++i) for (int k = 0; k < 3; ++k) c.palette_rgb[i][k] = colors.palette.list[i][k];
        c.bright_offset = colors.brighness.offset; c.bright_factor = colors.brighness.factor;
        return c;
    }
};

// ImageBuffer<Rgba<u16>, Vec<u16>> (lib.rs:625): interleaved RGBA, row-major.
struct FinalImage {
    uint32_t width = 0, height = 0;
    std::vector<uint16_t> raw;
    const uint16_t *pixel(uint32_t x, uint32_t y) const { return raw.data() + (size_t(y) * width + x) * 4; }
};

class Runtime {                                                    // lib.rs:631-739
public:
    // Runtime::new(&config).  seed: the reference seeds from the OS (lib.rs:656); pass one for reproducibility.
    explicit Runtime(const Config &config, int device = 0, uint64_t seed = std::random_device{}() | (uint64_t(std::random_device{}()) << 32))
        : seed_(seed) { check(sar_runtime_new(config.width, config.height, device, &h_)); }
    Runtime(const Runtime &) = delete;
    Runtime &operator=(const Runtime &) = delete;
    ~Runtime() { sar_runtime_free(h_); }
    void reset() { check(sar_runtime_reset(h_)); }                                   // lib.rs:682
    void merge(const Runtime &other) { check(sar_runtime_merge(h_, other.h_)); }     // lib.rs:708 (throws on dimension mismatch)
    sar_runtime *handle() const { return h_; }
    uint64_t seed_;
    uint64_t draws_ = 0;
private:
    sar_runtime *h_ = nullptr;
};

// render(&config, &mut runtime), lib.rs:747 — one trajectory, accumulates, does not reset.
inline void render(const Config &config, Runtime &runtime) {
    const sar_config c = config.to_pod();
    check(sar_render_seeded(&c, runtime.handle(), runtime.seed_, runtime.draws_, 1));
    ++runtime.draws_;
}
// the parity form: explicit start points (n×3 f64), one render() per point
inline void render(const Config &config, Runtime &runtime, const std::vector<double> &init_xyz) {
    const sar_config c = config.to_pod();
    check(sar_render(&c, runtime.handle(), init_xyz.data(), init_xyz.size() / 3));
}
// colorize(&config, &runtime) -> FinalImage, lib.rs:841
inline FinalImage colorize(const Config &config, const Runtime &runtime) {
    const sar_config c = config.to_pod();
    FinalImage img{config.width, config.height, std::vector<uint16_t>(size_t(config.width) * config.height * 4)};
    check(sar_colorize(&c, runtime.handle(), img.raw.data(), nullptr));
    return img;
}

class ParallelRenderer {                                           // lib.rs:908-1030
public:
    explicit ParallelRenderer(const std::vector<int> &devices = {}, uint32_t threads_per_device = 0) {
        check(sar_renderer_new(devices.empty() ? nullptr : devices.data(), int(devices.size()), threads_per_device, &h_));
    }
    ParallelRenderer(const ParallelRenderer &) = delete;
    ParallelRenderer &operator=(const ParallelRenderer &) = delete;
    ~ParallelRenderer() { shutdown(); }
    void shutdown() { sar_renderer_shutdown(h_); h_ = nullptr; }   // lib.rs:1020
    uint64_t num_threads() const { uint64_t n = 0; check(sar_renderer_num_threads(h_, &n)); return n; }   // lib.rs:1015
    sar_renderer *handle() const { return h_; }
private:
    sar_renderer *h_ = nullptr;
};

// render_parallel(&mut renderer, config, jobs_per_thread) -> FinalImage, lib.rs:1051
inline FinalImage render_parallel(ParallelRenderer &renderer, const Config &config, size_t jobs_per_thread,
                                  uint64_t seed = std::random_device{}() | (uint64_t(std::random_device{}()) << 32)) {
    const sar_config c = config.to_pod();
    FinalImage img{config.width, config.height, std::vector<uint16_t>(size_t(config.width) * config.height * 4)};
    check(sar_render_parallel(renderer.handle(), &c, jobs_per_thread, seed, nullptr, img.raw.data()));
    return img;
}

// ---- output conversion + raw encoders (src/bin/main.rs:40-100) ----------------------------------
enum class PixelFormat { Rgba16 = SAR_PIX_RGBA16, Rgb16 = SAR_PIX_RGB16, Rgba8 = SAR_PIX_RGBA8, Rgb8 = SAR_PIX_RGB8 };
enum class Container { Raw = SAR_FILE_RAW, Pam = SAR_FILE_PAM, Bmp = SAR_FILE_BMP, Png = SAR_FILE_PNG,
                       PngDeflate = SAR_FILE_PNG_DEFLATE /* frame sequences only; single images: encode_png() */ };
// the match at main.rs:52-57
inline PixelFormat pixel_format(bool transparent, bool eight_bit) {
    return transparent ? (eight_bit ? PixelFormat::Rgba8 : PixelFormat::Rgba16) : (eight_bit ? PixelFormat::Rgb8 : PixelFormat::Rgb16);
}
// The image of the last colorize() on `runtime`, converted on the device and wrapped in the container.
inline std::vector<uint8_t> encode_image(const Runtime &runtime, uint32_t width, uint32_t height, PixelFormat fmt, Container cont) {
    const size_t n = sar_encoded_size(width, height, uint32_t(fmt), uint32_t(cont));
    if (n == 0) throw Error(SAR_ERR_UNSUPPORTED, "this pixel format cannot be written in this container (BMP is 8-bit only)");
    std::vector<uint8_t> out(n);
    check(sar_runtime_encode(runtime.handle(), uint32_t(fmt), uint32_t(cont), out.data(), out.size(), nullptr));
    return out;
}
// The same image as a complete, deflate-compressed PNG (main.rs:78-89; compressor on the device, sar.h: sar_runtime_encode_png).
inline std::vector<uint8_t> encode_png(const Runtime &runtime, uint32_t width, uint32_t height, PixelFormat fmt) {
    const size_t cap = sar_png_bound(width, height, uint32_t(fmt));
    if (cap == 0) throw Error(SAR_ERR_UNSUPPORTED, "image too large for one IDAT chunk");
    std::vector<uint8_t> out(cap);
    size_t n = 0;
    check(sar_runtime_encode_png(runtime.handle(), uint32_t(fmt), out.data(), out.size(), &n, nullptr));
    out.resize(n);
    return out;
}
inline void write_image(const Runtime &runtime, const Config &config, const std::string &path, bool eight_bit, Container cont) {
    const auto bytes = cont == Container::Png ? encode_png(runtime, config.width, config.height, pixel_format(config.transparent, eight_bit))
                                              : encode_image(runtime, config.width, config.height, pixel_format(config.transparent, eight_bit), cont);
    check(sar_write_file(path.c_str(), bytes.data(), bytes.size()));
}

// ---- frame sequences: AngleIter and the binary's frame loop (src/bin/main.rs:107-176, 496-512) ------
// The angles AngleIter yields (main.rs:136-176): while curr + step/2 < end { yield curr (DEGREES) converted to radians,
// main.rs:166; curr += step }; if that is nothing, the single value `start` UNCONVERTED, like the reference's
// single-image branch (main.rs:168-170).  (A non-positive step, with which the reference never terminates, yields that too.)
inline std::vector<double> angle_iter(double start, double end, double step) {
    std::vector<double> out;
    const double pi = 3.14159265358979323846;
    if (step > 0.0)
        for (double curr = start; curr + step / 2. < end; curr += step) out.push_back(curr * pi / 180.);
    if (out.empty()) out.push_back(start);
    return out;
}
namespace detail {
template <class F> struct Thunk {
    static void frame16(void *user, uint32_t frame, const uint16_t *rgba) { (*static_cast<F *>(user))(frame, rgba); }
    static void bytes(void *user, uint32_t frame, const uint8_t *data, size_t n) { (*static_cast<F *>(user))(frame, data, n); }
};
}  // namespace detail
// for angle in angles { config.angle = angle; image = render_parallel(..); on_frame(index, rgba16 pixels) } — frames are
// rendered back to back on the renderer's devices, each frame's copy-out overlapping the next frame's render; on_frame is
// called in frame order from the calling thread (the pixels are only valid inside it).
template <class F>
inline void render_sequence(ParallelRenderer &renderer, const Config &config, const std::vector<double> &angles_rad,
                            size_t jobs_per_thread, uint64_t seed, F &&on_frame, bool shared_points = false) {
    const sar_config c = config.to_pod();
    using Fn = std::remove_reference_t<F>;
    check(sar_render_sequence(renderer.handle(), &c, angles_rad.data(), uint32_t(angles_rad.size()), jobs_per_thread, seed,
                              shared_points ? SAR_SEQ_SHARED_POINTS : 0u, nullptr, &detail::Thunk<Fn>::frame16, &on_frame));
}
// the same with every frame converted on the device and handed over encoded — what the reference's encoder side threads
// write (main.rs:508-511): on_frame(index, bytes, n_bytes).  Container::PngDeflate = complete compressed PNG files.
template <class F>
inline void render_sequence_encoded(ParallelRenderer &renderer, const Config &config, const std::vector<double> &angles_rad,
                                    size_t jobs_per_thread, uint64_t seed, PixelFormat fmt, Container cont, F &&on_frame,
                                    bool shared_points = false) {
    const sar_config c = config.to_pod();
    using Fn = std::remove_reference_t<F>;
    check(sar_render_sequence_encoded(renderer.handle(), &c, angles_rad.data(), uint32_t(angles_rad.size()), jobs_per_thread, seed,
                                      shared_points ? SAR_SEQ_SHARED_POINTS : 0u, uint32_t(fmt), uint32_t(cont), nullptr,
                                      &detail::Thunk<Fn>::bytes, &on_frame));
}

// ---- auto-framing first pass (lib.rs:326-334) ----------------------------------------------------
struct AutoFrame {
    std::array<double, 6> box;      // screen-space xmin, xmax, ymin, ymax, zmin, zmax (the table of lib.rs:329-333)
    Vec3 center_camera;
    double scale;
    uint64_t diverged, n_jobs;
    void apply(Config &config) const { config.view.center_camera = center_camera; config.view.scale = scale; }
};
inline AutoFrame autoframe(const Config &config, uint64_t n_jobs = 4096, uint64_t iterations = 20'000, uint64_t seed = 0, int device = 0) {
    const sar_config c = config.to_pod();
    sar_autoframe_result r;
    check(sar_autoframe(&c, device, seed, nullptr, n_jobs, iterations, &r));
    AutoFrame a;
    for (int i = 0; i < 6; ++i) a.box[i] = r.box[i];
    a.center_camera = {r.center_camera[0], r.center_camera[1], r.center_camera[2]};
    a.scale = r.scale; a.diverged = r.diverged; a.n_jobs = r.n_jobs;
    return a;
}

}  // namespace sar
