//! Drop-in replacements for the reference's `Runtime`, `render`, `colorize`, `ParallelRenderer`
//! and `render_parallel` (strange-attractor-renderer `src/lib.rs:631-1082`) backed by
//! `libsar_b200.so` (C ABI: `include/sar.h`).  SOURCE ONLY — the build image has no Rust toolchain,
//! so this file documents the binding a maintainer adds; it has not been compiled.
//!
//! Only the shipped instantiations can run on the GPU: `PolynomialSprott2Degree` with
//! `color_transforms::poisson_saturne` or `AdjustedVelocity`, wrapped in [`GpuConfig`] together with
//! the palette list (which the reference keeps private).  They implement [`DeviceConfig`]; any other
//! `Attractor` / `ColorTransform` keeps using the reference's CPU functions.
use std::ffi::CStr;
use std::os::raw::{c_char, c_int};

use strange_attractor_renderer as reference;
use reference::{attractors::PolynomialSprott2Degree, config::color_transforms, Config, RenderKind};

pub type FinalImage = reference::FinalImage;

pub const SAR_MAX_PALETTE: usize = 16;

/// `sar_config`, include/sar.h — the POD form of `Config` (lib.rs:265-287).
#[repr(C)]
#[derive(Clone, Copy)]
pub struct SarConfig {
    pub iterations: u64,
    pub width: u32,
    pub height: u32,
    pub render_kind: u32,
    pub transparent: u32,
    pub silent: u32,
    pub ct_kind: u32,
    pub angle: f64,
    pub coef: [[f64; 10]; 3],
    pub center_camera: [f64; 3],
    pub axis: [f64; 3],
    pub rotation: f64,
    pub scale: f64,
    pub ct_offset: f64,
    pub ct_factor: f64,
    pub palette_len: u32,
    pub attractor_kind: u32,
    pub palette_rgb: [[f64; 3]; SAR_MAX_PALETTE],
    pub bright_offset: f64,
    pub bright_factor: f64,
    pub coef3: [[f64; 10]; 3],
    pub ct_weights: [f64; 4],
}

#[repr(C)]
pub struct SarRuntime {
    _private: [u8; 0],
}
#[repr(C)]
pub struct SarRenderer {
    _private: [u8; 0],
}

extern "C" {
    fn sar_last_error() -> *const c_char;
    fn sar_runtime_new(width: u32, height: u32, device: c_int, out: *mut *mut SarRuntime) -> c_int;
    fn sar_runtime_free(rt: *mut SarRuntime);
    fn sar_runtime_reset(rt: *mut SarRuntime) -> c_int;
    fn sar_runtime_merge(dst: *mut SarRuntime, src: *const SarRuntime) -> c_int;
    fn sar_render_seeded(cfg: *const SarConfig, rt: *mut SarRuntime, seed: u64, first_job: u64, n_jobs: u64) -> c_int;
    fn sar_colorize(cfg: *const SarConfig, rt: *const SarRuntime, rgba_u16: *mut u16, rgba_f32: *mut f32) -> c_int;
    fn sar_renderer_new(devices: *const c_int, n_devices: c_int, threads_per_device: u32, out: *mut *mut SarRenderer) -> c_int;
    fn sar_renderer_shutdown(r: *mut SarRenderer);
    fn sar_renderer_runtime(r: *mut SarRenderer, out: *mut *mut SarRuntime) -> c_int;
    fn sar_png_bound(width: u32, height: u32, pixel_format: u32) -> usize;
    fn sar_runtime_encode_png(rt: *mut SarRuntime, pixel_format: u32, out: *mut u8, out_capacity: usize, out_bytes: *mut usize,
                              stream: *mut std::ffi::c_void) -> c_int;
    fn sar_render_parallel(r: *mut SarRenderer, cfg: *const SarConfig, jobs_per_thread: u64, seed: u64,
                           init_xyz: *const f64, rgba_u16: *mut u16) -> c_int;
}

/// The reference never returns `Result` from this path: it panics (lib.rs:678, 709-710, 1024).
/// Non-zero statuses become the same panics.
fn check(rc: c_int) {
    if rc != 0 {
        let msg = unsafe { CStr::from_ptr(sar_last_error()) }.to_string_lossy().into_owned();
        panic!("sar_b200 error {rc}: {msg}");
    }
}

fn os_seed() -> u64 {
    // stands in for SmallRng::from_os_rng() (lib.rs:656)
    use std::collections::hash_map::RandomState;
    use std::hash::{BuildHasher, Hasher};
    RandomState::new().build_hasher().finish()
}

/// Configs the GPU can run.  `palette` is the colour list `config.colors.palette` was built from
/// (`Palette::new` / `from_rgb`, lib.rs:413-433, WITHOUT the sentinel `new` appends): the reference keeps
/// that list private (lib.rs:408-411), so it has to be handed over explicitly — see [`GpuConfig`].
pub trait DeviceConfig {
    fn to_pod(&self) -> SarConfig;
}

/// The default palette, `Colors::default()` (lib.rs:483-487).
pub const DEFAULT_PALETTE: [[f64; 3]; 6] =
    [[1., 1., 0.5], [0.5, 1., 0.5], [1., 0.5, 0.5], [0.5, 1., 1.], [0.5, 0.5, 1.], [1., 0.5, 1.]];

/// A `Config` together with the palette list it was built from.  There is deliberately no way to
/// construct one without saying what the palette is: a custom `Palette` must never be replaced by
/// the default one silently (the GPU image would then differ from the reference's `colorize`).
pub struct GpuConfig<A: reference::Attractor, T: reference::ColorTransform> {
    pub config: Config<A, T>,
    palette: Vec<[f64; 3]>,
}
impl<A: reference::Attractor, T: reference::ColorTransform> GpuConfig<A, T> {
    /// `palette`: the list given to `Palette::new` / `from_rgb` for `config.colors.palette`.  Checked
    /// against the palette through its public `interpolate` (lib.rs:442) at the knots: panics on a mismatch.
    pub fn new(config: Config<A, T>, palette: &[[f64; 3]]) -> Self {
        assert!(!palette.is_empty() && palette.len() <= SAR_MAX_PALETTE, "1..=16 palette entries cross the C ABI");
        assert_eq!(config.colors.palette.count(), palette.len(), "palette list does not match config.colors.palette");
        for (k, rgb) in palette.iter().enumerate() {
            let got = config.colors.palette.interpolate(k as f64 / palette.len() as f64);
            for (g, want) in got.0.iter().zip(rgb.iter()) {
                assert!((g - want.sqrt()).abs() < 1e-12, "palette list does not match config.colors.palette at entry {k}");
            }
        }
        Self { config, palette: palette.to_vec() }
    }
    /// For configs that use `Colors::default()` (as `Config::new` does, lib.rs:303); verified like `new`.
    pub fn with_default_colors(config: Config<A, T>) -> Self {
        Self::new(config, &DEFAULT_PALETTE)
    }
}

fn fill_common<A: reference::Attractor, T: reference::ColorTransform>(
    c: &Config<A, T>, coeffs: &PolynomialSprott2Degree, ct_kind: u32, ct_offset: f64, ct_factor: f64,
    palette: &[[f64; 3]],
) -> SarConfig {
    let mut palette_rgb = [[0.0; 3]; SAR_MAX_PALETTE];
    palette_rgb[..palette.len()].copy_from_slice(palette);
    SarConfig {
        iterations: c.iterations as u64,
        width: c.width,
        height: c.height,
        render_kind: match c.render { RenderKind::Gas => 0, RenderKind::Depth => 1 },
        transparent: c.transparent as u32,
        silent: c.silent as u32,
        ct_kind,
        angle: c.angle,
        coef: [coeffs.x, coeffs.y, coeffs.z],
        center_camera: [c.view.center_camera.x, c.view.center_camera.y, c.view.center_camera.z],
        axis: [c.view.rotation.axis.x, c.view.rotation.axis.y, c.view.rotation.axis.z],
        rotation: c.view.rotation.rotation,
        scale: c.view.scale,
        ct_offset,
        ct_factor,
        palette_len: palette.len() as u32,
        attractor_kind: 0, // PolynomialSprott2Degree, the only Attractor the reference ships
        palette_rgb,
        bright_offset: c.colors.brighness.offset,
        bright_factor: c.colors.brighness.factor,
        coef3: [[0.0; 10]; 3],
        ct_weights: [0.0; 4],
    }
}

impl DeviceConfig for GpuConfig<PolynomialSprott2Degree, color_transforms::AdjustedVelocity> {
    fn to_pod(&self) -> SarConfig {
        let c = &self.config;
        fill_common(c, &c.attractor, 1, c.color_transform.offset, c.color_transform.factor, &self.palette)
    }
}
impl DeviceConfig for GpuConfig<PolynomialSprott2Degree, color_transforms::Function> {
    /// Only valid when `color_transform` is `color_transforms::poisson_saturne`; a different fn
    /// pointer has no device form and must stay on the reference's CPU path.
    fn to_pod(&self) -> SarConfig {
        let c = &self.config;
        assert!(c.color_transform as usize == color_transforms::poisson_saturne as usize,
                "only color_transforms::poisson_saturne has a device implementation");
        fill_common(c, &c.attractor, 0, 0.0, 0.0, &self.palette)
    }
}

/// `Runtime` (lib.rs:631-646) living in GPU memory.
pub struct Runtime {
    handle: *mut SarRuntime,
    seed: u64,
    draws: u64,
}
unsafe impl Send for Runtime {}

impl Runtime {
    /// `Runtime::new(&config)`, lib.rs:660.
    pub fn new(config: &impl DeviceConfig) -> Self {
        let pod = config.to_pod();
        let mut handle = std::ptr::null_mut();
        check(unsafe { sar_runtime_new(pod.width, pod.height, 0, &mut handle) });
        Self { handle, seed: os_seed(), draws: 0 }
    }
    /// `Runtime::reset`, lib.rs:682.
    pub fn reset(&mut self) {
        check(unsafe { sar_runtime_reset(self.handle) });
    }
    /// `Runtime::merge`, lib.rs:708.  Panics on a dimension mismatch like the reference.
    pub fn merge(&mut self, other: &Self) {
        check(unsafe { sar_runtime_merge(self.handle, other.handle) });
    }
}
impl Drop for Runtime {
    fn drop(&mut self) {
        unsafe { sar_runtime_free(self.handle) }
    }
}

/// `render(&config, &mut runtime)`, lib.rs:747.
pub fn render(config: &impl DeviceConfig, runtime: &mut Runtime) {
    let pod = config.to_pod();
    check(unsafe { sar_render_seeded(&pod, runtime.handle, runtime.seed, runtime.draws, 1) });
    runtime.draws += 1;
}

/// `colorize(&config, &runtime) -> FinalImage`, lib.rs:841.
#[must_use]
pub fn colorize(config: &impl DeviceConfig, runtime: &Runtime) -> FinalImage {
    let pod = config.to_pod();
    let mut raw = vec![0u16; pod.width as usize * pod.height as usize * 4];
    check(unsafe { sar_colorize(&pod, runtime.handle, raw.as_mut_ptr(), std::ptr::null_mut()) });
    image::ImageBuffer::from_raw(pod.width, pod.height, raw).unwrap()
}

/// The default branch of `write_image_matches` (src/bin/main.rs:52-57, 78-89) for the image of the last `colorize` /
/// `render_parallel` on this runtime: `(transparent, eight_bit)` picks RGBA16 / RGB16 / RGBA8 / RGB8, and the complete PNG
/// file — filtered, deflated and checksummed on the device — comes back as bytes to hand to `File::write_all`.
#[must_use]
pub fn encode_png(runtime: &Runtime, width: u32, height: u32, transparent: bool, eight_bit: bool) -> Vec<u8> {
    let fmt = match (transparent, eight_bit) {
        (true, false) => 0u32,
        (false, false) => 1,
        (true, true) => 2,
        (false, true) => 3,
    };
    let cap = unsafe { sar_png_bound(width, height, fmt) };
    assert!(cap != 0, "image too large for one IDAT chunk");
    let mut out = vec![0u8; cap];
    let mut n = 0usize;
    check(unsafe { sar_runtime_encode_png(runtime.handle, fmt, out.as_mut_ptr(), cap, &mut n, std::ptr::null_mut()) });
    out.truncate(n);
    out
}

/// `ParallelRenderer` (lib.rs:908-915): the worker threads are the GPU's trajectory lanes.
pub struct ParallelRenderer {
    handle: *mut SarRenderer,
}
impl ParallelRenderer {
    /// `ParallelRenderer::new()`, lib.rs:919.
    pub fn new() -> Self {
        let mut handle = std::ptr::null_mut();
        check(unsafe { sar_renderer_new(std::ptr::null(), 0, 0, &mut handle) });
        Self { handle }
    }
    /// `shutdown(self)`, lib.rs:1020.
    pub fn shutdown(self) {}
    /// The merged Runtime of the last `render_parallel` (borrowed from the renderer): what `encode_png` reads.
    pub fn with_runtime<R>(&mut self, f: impl FnOnce(&Runtime) -> R) -> R {
        let mut handle = std::ptr::null_mut();
        check(unsafe { sar_renderer_runtime(self.handle, &mut handle) });
        let borrowed = std::mem::ManuallyDrop::new(Runtime { handle, seed: 0, draws: 0 });
        f(&borrowed)
    }
}
impl Default for ParallelRenderer {
    fn default() -> Self {
        Self::new()
    }
}
impl Drop for ParallelRenderer {
    fn drop(&mut self) {
        unsafe { sar_renderer_shutdown(self.handle) }
    }
}

/// `render_parallel(&mut renderer, config, jobs_per_thread) -> FinalImage`, lib.rs:1051.
#[must_use]
pub fn render_parallel(renderer: &mut ParallelRenderer, config: impl DeviceConfig, jobs_per_thread: usize) -> FinalImage {
    let pod = config.to_pod();
    let mut raw = vec![0u16; pod.width as usize * pod.height as usize * 4];
    check(unsafe {
        sar_render_parallel(renderer.handle, &pod, jobs_per_thread as u64, os_seed(), std::ptr::null(), raw.as_mut_ptr())
    });
    image::ImageBuffer::from_raw(pod.width, pod.height, raw).unwrap()
}
