"""Host-side mirror of the reference's public API for the render path, over the C ABI.

Same names, argument meaning and error behaviour as `strange_attractor_renderer` (src/lib.rs):

    Config / View / Colors / Palette / BrighnessConstants / RenderKind      lib.rs:232-492
    attractors.PolynomialSprott2Degree                                      lib.rs:575-580
    color_transforms.{poisson_saturne, AdjustedVelocity}                    lib.rs:503-559
    Runtime.{new, reset, merge}                                             lib.rs:660, 682, 708
    render(config, runtime)                                                 lib.rs:747
    colorize(config, runtime) -> FinalImage                                 lib.rs:841
    ParallelRenderer.{new, shutdown}, render_parallel(...)                  lib.rs:919, 1020, 1051

Everything that computes runs in libsar_b200.so on the GPU; this file only marshals.
Where the reference panics (dimension mismatch in merge, lib.rs:709-710; empty palette,
lib.rs:415-418) a `SarError` is raised instead.

One deliberate extension: the reference draws start points from an OS-seeded SmallRng
(lib.rs:656, 748) and is therefore not reproducible.  `Runtime.new(config, seed=...)` and
`render_parallel(..., seed=...)` / `initial_points=` make the start points explicit; with no
seed given an OS-random one is used, as in the reference.
"""
from __future__ import annotations

import ctypes as C
import enum
import os
import warnings
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import _native as N
from ._native import SarConfig, SarError  # noqa: F401


# ---- primitives (lib.rs:79-224) ---------------------------------------------------------
@dataclass
class Vec3:
    x: float
    y: float
    z: float

    @staticmethod
    def new(x: float, y: float, z: float) -> "Vec3":
        return Vec3(x, y, z)


@dataclass
class EulerAxisRotation:
    """lib.rs:170-175.  NB the axis is used as given — the reference normalises it only in
    debug builds (lib.rs:181-183) and its published images come from release builds."""

    axis: Vec3
    rotation: float


# ---- config (lib.rs:226-560) ------------------------------------------------------------
class RenderKind(enum.Enum):
    Gas = N.SAR_RENDER_GAS
    Depth = N.SAR_RENDER_DEPTH


@dataclass
class View:
    center_camera: Vec3
    rotation: EulerAxisRotation
    scale: float


@dataclass
class BrighnessConstants:  # sic, lib.rs:390
    offset: float = -0.15
    factor: float = 5.0 / 3.0


class Palette:
    """lib.rs:408-473.  `list` holds the colours WITHOUT the duplicated sentinel the reference
    appends (lib.rs:418); the library re-creates it."""

    def __init__(self, colors: Sequence[Sequence[float]]):
        if len(colors) == 0:
            raise SarError(N.SAR_ERR_INVALID, "Palette::new panics if list.is_empty() (lib.rs:415)")
        if len(colors) > N.SAR_MAX_PALETTE:
            raise SarError(N.SAR_ERR_INVALID, f"at most {N.SAR_MAX_PALETTE} palette entries cross the C ABI")
        self.list = [tuple(float(v) for v in c) for c in colors]

    @staticmethod
    def new(colors):
        return Palette(colors)

    @staticmethod
    def from_rgb(r, g, b) -> "Palette":
        return Palette(list(zip(r, g, b)))

    def count(self) -> int:
        return len(self.list)


@dataclass
class Colors:
    palette: Palette = field(default_factory=lambda: Palette.from_rgb(
        [1.0, 0.5, 1.0, 0.5, 0.5, 1.0], [1.0, 1.0, 0.5, 1.0, 0.5, 0.5], [0.5, 0.5, 0.5, 1.0, 1.0, 1.0]))
    brighness: BrighnessConstants = field(default_factory=BrighnessConstants)


class attractors:  # namespace, lib.rs:567
    @dataclass
    class PolynomialSprott2Degree:
        x: List[float]
        y: List[float]
        z: List[float]


    @dataclass
    class PolynomialSprott3Degree:
        """The cubic member of the family (include/sar.h: SAR_ATTRACTOR_SPROTT3) — an Attractor the
        reference does not ship (README.md:8).  x/y/z: the 10 quadratic coefficients in the reference's
        order; x3/y3/z3: those of [x³, x²y, x²z, xy², xyz, xz², y³, y²z, yz², z³]."""

        x: List[float]
        y: List[float]
        z: List[float]
        x3: List[float]
        y3: List[float]
        z3: List[float]


class color_transforms:  # namespace, lib.rs:498
    @dataclass
    class AdjustedVelocity:
        offset: float
        factor: float

    @dataclass
    class ScreenBlend:
        """A device form of the closures ColorTransform also accepts (lib.rs:245; include/sar.h:
        SAR_CT_SCREEN_BLEND): ((s.x*w0 + s.y*w1 + s.z*w2 + |delta|*w3) + offset) * factor."""

        weights: List[float]
        offset: float = 0.0
        factor: float = 1.0

    class _PoissonSaturne:
        """Marker for the fn item `color_transforms::poisson_saturne` (lib.rs:520)."""

        def __repr__(self):
            return "color_transforms.poisson_saturne"

    poisson_saturne = _PoissonSaturne()


@dataclass
class Config:
    """lib.rs:265-287; defaults of Config::new (lib.rs:289-307)."""

    attractor: "attractors.PolynomialSprott2Degree"
    view: View
    color_transform: object
    iterations: int = 10_000_000
    width: int = 1920
    height: int = 1080
    render: RenderKind = RenderKind.Gas
    transparent: bool = True
    angle: float = 0.0
    silent: bool = True
    colors: Colors = field(default_factory=Colors)

    @staticmethod
    def new(coefficients, view, transform_colors) -> "Config":
        return Config(coefficients, view, transform_colors)

    @staticmethod
    def _from_pod(c: SarConfig) -> "Config":
        if c.attractor_kind == N.SAR_ATTRACTOR_SPROTT3:
            att = attractors.PolynomialSprott3Degree(list(c.coef[0]), list(c.coef[1]), list(c.coef[2]),
                                                     list(c.coef3[0]), list(c.coef3[1]), list(c.coef3[2]))
        else:
            att = attractors.PolynomialSprott2Degree(list(c.coef[0]), list(c.coef[1]), list(c.coef[2]))
        view = View(Vec3(*c.center_camera), EulerAxisRotation(Vec3(*c.axis), c.rotation), c.scale)
        ct = (color_transforms.poisson_saturne if c.ct_kind == N.SAR_CT_POISSON_SATURNE
              else color_transforms.AdjustedVelocity(offset=c.ct_offset, factor=c.ct_factor) if c.ct_kind == N.SAR_CT_ADJUSTED_VELOCITY
              else color_transforms.ScreenBlend(list(c.ct_weights), offset=c.ct_offset, factor=c.ct_factor))
        pal = Palette([tuple(c.palette_rgb[i]) for i in range(c.palette_len)])
        return Config(att, view, ct, iterations=c.iterations, width=c.width, height=c.height,
                      render=RenderKind(c.render_kind), transparent=bool(c.transparent), angle=c.angle,
                      silent=bool(c.silent), colors=Colors(pal, BrighnessConstants(c.bright_offset, c.bright_factor)))

    @staticmethod
    def poisson_saturne() -> "Config":
        """Config::poisson_saturne(), lib.rs:310-352 (constants held by the library)."""
        c = SarConfig()
        N.check(N.lib().sar_config_poisson_saturne(C.byref(c)))
        return Config._from_pod(c)

    @staticmethod
    def solar_sail() -> "Config":
        """Config::solar_sail(), lib.rs:355-386."""
        c = SarConfig()
        N.check(N.lib().sar_config_solar_sail(C.byref(c)))
        return Config._from_pod(c)

    def to_pod(self) -> SarConfig:
        c = SarConfig()
        c.iterations, c.width, c.height = int(self.iterations), int(self.width), int(self.height)
        c.render_kind = self.render.value
        c.transparent, c.silent, c.angle = int(bool(self.transparent)), int(bool(self.silent)), float(self.angle)
        a = self.attractor
        if isinstance(a, attractors.PolynomialSprott3Degree):
            c.attractor_kind = N.SAR_ATTRACTOR_SPROTT3
            for k, lst in enumerate((a.x3, a.y3, a.z3)):
                if len(lst) != 10:
                    raise SarError(N.SAR_ERR_INVALID, "cubic coefficient lists have 10 entries (include/sar.h)")
                for i, v in enumerate(lst):
                    c.coef3[k][i] = float(v)
        elif not isinstance(a, attractors.PolynomialSprott2Degree):
            raise SarError(N.SAR_ERR_UNSUPPORTED, "only PolynomialSprott2Degree / PolynomialSprott3Degree have a device implementation")
        for k, lst in enumerate((a.x, a.y, a.z)):
            if len(lst) != 10:
                raise SarError(N.SAR_ERR_INVALID, "coefficient lists have 10 entries (lib.rs:577-579)")
            for i, v in enumerate(lst):
                c.coef[k][i] = float(v)
        v = self.view
        c.center_camera[:] = [v.center_camera.x, v.center_camera.y, v.center_camera.z]
        c.axis[:] = [v.rotation.axis.x, v.rotation.axis.y, v.rotation.axis.z]
        c.rotation, c.scale = float(v.rotation.rotation), float(v.scale)
        t = self.color_transform
        if isinstance(t, color_transforms.AdjustedVelocity):
            c.ct_kind, c.ct_offset, c.ct_factor = N.SAR_CT_ADJUSTED_VELOCITY, float(t.offset), float(t.factor)
        elif isinstance(t, color_transforms.ScreenBlend):
            if len(t.weights) != 4:
                raise SarError(N.SAR_ERR_INVALID, "ScreenBlend takes 4 weights (include/sar.h)")
            c.ct_kind, c.ct_offset, c.ct_factor = N.SAR_CT_SCREEN_BLEND, float(t.offset), float(t.factor)
            c.ct_weights[:] = [float(v) for v in t.weights]
        elif t is color_transforms.poisson_saturne:
            c.ct_kind = N.SAR_CT_POISSON_SATURNE
        else:
            raise SarError(N.SAR_ERR_UNSUPPORTED,
                           "only color_transforms.poisson_saturne, AdjustedVelocity and ScreenBlend have a device implementation")
        pal = self.colors.palette
        c.palette_len = pal.count()
        for i, rgb in enumerate(pal.list):
            c.palette_rgb[i][:] = list(rgb)
        c.bright_offset, c.bright_factor = float(self.colors.brighness.offset), float(self.colors.brighness.factor)
        return c


def _pod(config) -> SarConfig:
    return config if isinstance(config, SarConfig) else config.to_pod()


def seed_points(seed: int, first: int, n: int) -> np.ndarray:
    """The documented start-point generator (include/sar.h: sar_seed_points), [n,3] f64."""
    out = np.empty((n, 3), dtype=np.float64)
    N.check(N.lib().sar_seed_points(seed & (2**64 - 1), first, n, out.ctypes.data_as(N._f64p)))
    return out


_warned_serial = False
FinalImage = np.ndarray  # [height, width, 4] uint16, RGBA — ImageBuffer<Rgba<u16>, Vec<u16>>, lib.rs:625


# ---- Runtime (lib.rs:631-739) -----------------------------------------------------------
class Runtime:
    """Device-resident count / steps / zbuf / max.  Construct with Runtime.new(config)."""

    def __init__(self, handle, width, height, device, seed, owned=True):
        self._h, self.width, self.height, self.device = handle, width, height, device
        self._seed = seed
        self._draws = 0   # how many start points this Runtime's generator has handed out
        self._owned = owned

    @staticmethod
    def new(config, device: int = 0, seed: Optional[int] = None) -> "Runtime":
        c = _pod(config)
        h = C.c_void_p()
        N.check(N.lib().sar_runtime_new(c.width, c.height, device, C.byref(h)))
        if seed is None:
            seed = int.from_bytes(os.urandom(8), "little")   # SmallRng::from_os_rng(), lib.rs:656
        return Runtime(h, c.width, c.height, device, seed)

    def close(self) -> None:
        if getattr(self, "_h", None) and self._owned:
            N.lib().sar_runtime_free(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self) -> None:
        N.check(N.lib().sar_runtime_reset(self._h))

    def merge(self, other: "Runtime") -> None:
        N.check(N.lib().sar_runtime_merge(self._h, other._h))

    # inspection (the reference keeps these fields private; tests and checkpoints need them)
    def download(self):
        """-> (count u32 [H,W], steps f64 [H,W], zbuf f32 [H,W], max)"""
        count = np.empty((self.height, self.width), dtype=np.uint32)
        steps = np.empty((self.height, self.width), dtype=np.float64)
        zbuf = np.empty((self.height, self.width), dtype=np.float32)
        mx = C.c_uint32()
        N.check(N.lib().sar_runtime_download(self._h, count.ctypes.data_as(N._u32p), steps.ctypes.data_as(N._f64p),
                                             zbuf.ctypes.data_as(N._f32p), C.byref(mx)))
        return count, steps, zbuf, int(mx.value)

    def upload(self, count, steps, zbuf) -> None:
        count = np.ascontiguousarray(count, dtype=np.uint32)
        steps = np.ascontiguousarray(steps, dtype=np.float64)
        zbuf = np.ascontiguousarray(zbuf, dtype=np.float32)
        if count.shape != (self.height, self.width) or steps.shape != count.shape or zbuf.shape != count.shape:
            raise SarError(N.SAR_ERR_DIMS, "upload: array shapes must be [height, width]")
        N.check(N.lib().sar_runtime_upload(self._h, count.ctypes.data_as(N._u32p), steps.ctypes.data_as(N._f64p),
                                           zbuf.ctypes.data_as(N._f32p)))


def render(config, runtime: Runtime, initial_points=None) -> None:
    """render(&config, &mut runtime), lib.rs:747: one trajectory of config.iterations recorded
    steps accumulated into `runtime` (which is NOT reset).  NB one trajectory is strictly serial
    (lib.rs:769-837) and occupies ONE GPU lane: this call is the parity interface, not the fast one —
    render_parallel() is what decomposes a frame over the device's ~130 000 lanes (lib.rs:1058-1062).  The start point comes from the
    Runtime's generator (seeded at Runtime.new) unless `initial_points` ([n,3] f64) is given, in
    which case it is n successive render() calls, one per point."""
    c = _pod(config)
    n_pts = 1 if initial_points is None else int(np.asarray(initial_points).size // 3)
    if c.iterations >= 1_000_000 and n_pts < 1024:
        # the reference's render() is ONE serial trajectory (lib.rs:769-837): on a GPU that is one lane out of
        # ~130 000, slower than a CPU core.  Exact, but almost never what a port wants: say so once per process.
        global _warned_serial
        if not _warned_serial:
            _warned_serial = True
            warnings.warn("render() runs each start point as one serial trajectory on ONE GPU lane (reference semantics, "
                          "lib.rs:747); for throughput use render_parallel() or pass >= 1e4 initial_points", stacklevel=2)
    if initial_points is None:
        N.check(N.lib().sar_render_seeded(C.byref(c), runtime._h, runtime._seed & (2**64 - 1), runtime._draws, 1))
        runtime._draws += 1
        return
    pts = np.ascontiguousarray(initial_points, dtype=np.float64).reshape(-1, 3)
    N.check(N.lib().sar_render(C.byref(c), runtime._h, pts.ctypes.data_as(N._f64p), pts.shape[0]))


def colorize(config, runtime: Runtime, want_f32: bool = False):
    """colorize(&config, &runtime) -> FinalImage, lib.rs:841.  want_f32 additionally returns the
    pre-quantisation colour buffer ([H,W,4] f32)."""
    c = _pod(config)
    out = np.empty((runtime.height, runtime.width, 4), dtype=np.uint16)
    f = np.empty((runtime.height, runtime.width, 4), dtype=np.float32) if want_f32 else None
    N.check(N.lib().sar_colorize(C.byref(c), runtime._h, out.ctypes.data_as(N._u16p),
                                 f.ctypes.data_as(N._f32p) if want_f32 else None))
    return (out, f) if want_f32 else out


# ---- ParallelRenderer / render_parallel (lib.rs:906-1082) -------------------------------
class ParallelRenderer:
    """`threads` plays available_parallelism() (lib.rs:920): concurrent trajectory lanes per
    device; 0 = library default (sar_default_threads: SM count × 896)."""

    def __init__(self, devices: Optional[Sequence[int]] = None, threads: int = 0):
        self._h = C.c_void_p()
        if devices:
            arr = (C.c_int * len(devices))(*devices)
            N.check(N.lib().sar_renderer_new(arr, len(devices), threads, C.byref(self._h)))
        else:
            N.check(N.lib().sar_renderer_new(None, 0, threads, C.byref(self._h)))

    @staticmethod
    def new(devices=None, threads: int = 0) -> "ParallelRenderer":
        return ParallelRenderer(devices, threads)

    default = new

    def num_threads(self, jobs_per_thread: int = 1) -> int:
        """lib.rs:1015.  In auto mode (threads=0) the count adapts to jobs_per_thread so that the
        number of jobs stays at the device's lane count (include/sar.h)."""
        n = C.c_uint64()
        N.check(N.lib().sar_renderer_num_threads_for(self._h, jobs_per_thread, C.byref(n)))
        return int(n.value)

    def plan(self, iterations: int, jobs_per_thread: int = 1):
        """(num_threads, iterations_per_job) render_parallel will use for a frame of `iterations`
        (include/sar.h: sar_renderer_plan; lib.rs:1058-1062)."""
        n, ipj = C.c_uint64(), C.c_uint64()
        N.check(N.lib().sar_renderer_plan(self._h, int(iterations), jobs_per_thread, C.byref(n), C.byref(ipj)))
        return int(n.value), int(ipj.value)

    def shutdown(self) -> None:
        if getattr(self, "_h", None):
            N.lib().sar_renderer_shutdown(self._h)
        self._h = None

    def __del__(self):
        try:
            self.shutdown()
        except Exception:
            pass

    def runtime(self) -> Runtime:
        """The merged Runtime of the last render_parallel (borrowed; valid until the next call)."""
        h = C.c_void_p()
        N.check(N.lib().sar_renderer_runtime(self._h, C.byref(h)))
        w, hh, dev = C.c_uint32(), C.c_uint32(), C.c_int()
        N.check(N.lib().sar_runtime_dims(h, C.byref(w), C.byref(hh), C.byref(dev)))
        return Runtime(h, int(w.value), int(hh.value), int(dev.value), 0, owned=False)


def render_parallel(renderer: ParallelRenderer, config, jobs_per_thread: int, seed: Optional[int] = None,
                    initial_points=None, out: Optional[np.ndarray] = None) -> FinalImage:
    """render_parallel(&mut renderer, config, jobs_per_thread) -> FinalImage, lib.rs:1051.
    num_threads*jobs_per_thread jobs of iterations/num_threads/jobs_per_thread steps each."""
    c = _pod(config)
    if seed is None:
        seed = int.from_bytes(os.urandom(8), "little")
    if out is None:
        out = np.empty((c.height, c.width, 4), dtype=np.uint16)
    pts_p = None
    if initial_points is not None:
        pts = np.ascontiguousarray(initial_points, dtype=np.float64).reshape(-1, 3)
        if pts.shape[0] < renderer.plan(c.iterations, jobs_per_thread)[0] * jobs_per_thread:
            raise SarError(N.SAR_ERR_INVALID, "initial_points must hold num_threads*jobs_per_thread points")
        pts_p = pts.ctypes.data_as(N._f64p)
    N.check(N.lib().sar_render_parallel(renderer._h, C.byref(c), jobs_per_thread, seed & (2**64 - 1), pts_p,
                                        out.ctypes.data_as(N._u16p)))
    return out


# ---- auto-framing first pass (the TODO at lib.rs:326-334) ----------------------------------
@dataclass
class AutoFrame:
    """box = screen-space (R·p) xmin, xmax, ymin, ymax, zmin, zmax of the bounded trajectories — the
    table of the comment at lib.rs:329-333; center_camera / scale = a View that centres and fits it."""

    box: List[float]
    center_camera: Vec3
    scale: float
    diverged: int
    n_jobs: int

    @property
    def diverged_fraction(self) -> float:
        return self.diverged / self.n_jobs if self.n_jobs else 0.0

    def apply(self, config: "Config") -> "Config":
        """config.view.center_camera / scale replaced by the derived ones (in place; returns config)."""
        config.view.center_camera = Vec3(self.center_camera.x, self.center_camera.y, self.center_camera.z)
        config.view.scale = self.scale
        return config


def autoframe(config, n_jobs: int = 4096, iterations: int = 20_000, seed: int = 0, initial_points=None, device: int = 0) -> AutoFrame:
    """The first pass the reference's author asks for (lib.rs:326-334), on the GPU: bounding box of the
    attractor in screen space from a batch of short trajectories, and the View values derived from it."""
    c = _pod(config)
    res = N.SarAutoframeResult()
    pts_p = None
    if initial_points is not None:
        pts = np.ascontiguousarray(initial_points, dtype=np.float64).reshape(-1, 3)
        n_jobs, pts_p = pts.shape[0], pts.ctypes.data_as(N._f64p)
    N.check(N.lib().sar_autoframe(C.byref(c), device, seed & (2**64 - 1), pts_p, n_jobs, iterations, C.byref(res)))
    return AutoFrame(list(res.box), Vec3(*res.center_camera), float(res.scale), int(res.diverged), int(res.n_jobs))


# ---- output conversion + raw encoders (src/bin/main.rs:40-100) ---------------------------
class PixelFormat(enum.Enum):
    """What write_image_matches converts the FinalImage to, by (transparent, 8bit) — main.rs:52-57."""

    Rgba16 = N.SAR_PIX_RGBA16   # (true, false): as is
    Rgb16 = N.SAR_PIX_RGB16     # (false, false): to_rgb16()
    Rgba8 = N.SAR_PIX_RGBA8     # (true, true): to_rgba8()
    Rgb8 = N.SAR_PIX_RGB8       # (false, true): to_rgb8()

    @staticmethod
    def of(transparent: bool, eight_bit: bool) -> "PixelFormat":
        return {(True, False): PixelFormat.Rgba16, (False, False): PixelFormat.Rgb16,
                (True, True): PixelFormat.Rgba8, (False, True): PixelFormat.Rgb8}[(bool(transparent), bool(eight_bit))]


class Container(enum.Enum):
    Raw = N.SAR_FILE_RAW        # the converted image.as_bytes() alone (feed it to a PNG encoder)
    Pam = N.SAR_FILE_PAM        # --pam, main.rs:62-68
    Bmp = N.SAR_FILE_BMP        # --bmp, main.rs:70-76 (8-bit formats only)
    Png = N.SAR_FILE_PNG        # the default branch, main.rs:78-89, without the compressor (stored deflate blocks)
    PngDeflate = N.SAR_FILE_PNG_DEFLATE   # the same with the compressor (device deflate); frame sequences only — see encode_png()


def encode_image(runtime: Runtime, pixel_format: PixelFormat, container: Container = Container.Raw) -> np.ndarray:
    """The image of the last colorize() on `runtime`, converted on the device and wrapped in the
    container: uint8 array of exactly the bytes the reference's encoder would write."""
    n = N.lib().sar_encoded_size(runtime.width, runtime.height, pixel_format.value, container.value)
    if n == 0:
        raise SarError(N.SAR_ERR_UNSUPPORTED, f"{pixel_format.name} cannot be written as {container.name}")
    out = np.empty(n, dtype=np.uint8)
    N.check(N.lib().sar_runtime_encode(runtime._h, pixel_format.value, container.value, out.ctypes.data_as(N._u8p), n, None))
    return out


def encode_png(runtime: Runtime, pixel_format: PixelFormat) -> np.ndarray:
    """The image of the last colorize() on `runtime` as a complete COMPRESSED PNG (main.rs:78-89): filter, deflate and
    checksums on the device (include/sar.h: sar_runtime_encode_png).  uint8 array of the file's bytes."""
    cap = N.lib().sar_png_bound(runtime.width, runtime.height, pixel_format.value)
    if cap == 0:
        raise SarError(N.SAR_ERR_UNSUPPORTED, "image too large for one IDAT chunk")
    out = np.empty(cap, dtype=np.uint8)
    n = C.c_size_t()
    N.check(N.lib().sar_runtime_encode_png(runtime._h, pixel_format.value, out.ctypes.data_as(N._u8p), cap, C.byref(n), None))
    return out[:n.value].copy()


def write_image(runtime: Runtime, path: str, transparent: bool, eight_bit: bool, container: Container, compress: bool = True) -> str:
    """write_image_matches (main.rs:40-100): convert, set the extension, write.  The PNG branch (main.rs:78-89) writes a
    deflate-compressed file like the reference's encoder unless compress=False (stored blocks, sar_runtime_encode)."""
    if container is Container.Png and compress:
        data = encode_png(runtime, PixelFormat.of(transparent, eight_bit))
    else:
        data = encode_image(runtime, PixelFormat.of(transparent, eight_bit), container)
    ext = {Container.Pam: ".pam", Container.Bmp: ".bmp", Container.Raw: ".raw", Container.Png: ".png"}[container]
    path = os.path.splitext(path)[0] + ext                       # name.set_extension(..), main.rs:63, 71
    N.check(N.lib().sar_write_file(path.encode(), data.ctypes.data_as(N._u8p), data.size))
    return path


# ---- frame sequences (src/bin/main.rs:107-176, 459-517) ---------------------------------
def angle_iter(start: float, end: float, step: float) -> List[float]:
    """The angles AngleIter yields (main.rs:107-176): while curr + step/2 < end, curr (DEGREES)
    converted to radians (main.rs:166), curr += step.  If that yields nothing, the single value
    `start` is returned unconverted, exactly like the reference's single-image branch
    (main.rs:169-171) — i.e. it is then taken as radians."""
    import math

    out, curr = [], float(start)
    while curr + step / 2.0 < end:
        out.append(curr * math.pi / 180.0)
        curr += step
    return out if out else [float(start)]


def angle_iter_files(start: float, end: float, step: float, file: str) -> List[tuple]:
    """AngleIter as the binary uses it (main.rs:106-176): (angle, path) pairs.  Frame i of a sequence is written to
    `<stem><i zero-padded to needed_digits><.ext>` next to `file`, with needed_digits = ceil(log10((end - start - step/2) / step))
    (0 when that count is <= 1, main.rs:116-122); a range that yields no frame gives the single pair (start, file)."""
    import math

    count = (end - start - step / 2.0) / step if step != 0 else 0.0
    as_usize = 0 if not (count > 0.0) else int(min(count, float(2**64 - 1)))          # `count as usize`: saturating, NaN -> 0
    digits = 0 if as_usize <= 1 else int(math.ceil(math.log10(count)))
    folder, name = os.path.split(file)
    stem, ext = os.path.splitext(name)                          # file_stem() / extension(), main.rs:141-156
    stem = stem or "attractor"
    out = []
    for i, a in enumerate(angle_iter(start, end, step)):
        if i == 0 and not (start + step / 2.0 < end):
            return [(a, file)]                                  # the single-image branch, main.rs:168-170
        idx = f"{i:0>{digits}}" if digits > 0 else ""
        name = stem + idx
        if ext:                                                 # PathBuf::set_extension: replaces what follows the last '.'
            cut = name.rfind(".")
            name = (name[:cut] if cut > 0 else name) + ext
        out.append((a, os.path.join(folder, name)))
    return out


def render_sequence(renderer: ParallelRenderer, config, angles: Sequence[float], jobs_per_thread: int,
                    seed: Optional[int] = None, shared_points: bool = False, out: Optional[np.ndarray] = None,
                    callback=None) -> Optional[np.ndarray]:
    """The binary's frame loop (main.rs:496-512): for each angle (radians) set config.angle and
    render_parallel.  Returns [n_frames, H, W, 4] uint16 (or streams frames to
    callback(frame_index, image_view) when given and out is None).  shared_points=True uses one
    list of start points for every frame, so the 1000-step warm-up runs once (include/sar.h)."""
    c = _pod(config)
    if seed is None:
        seed = int.from_bytes(os.urandom(8), "little")
    ang = np.ascontiguousarray(angles, dtype=np.float64)
    n = int(ang.shape[0])
    if out is None and callback is None:
        out = np.empty((n, c.height, c.width, 4), dtype=np.uint16)
    cb_c = None
    if callback is not None:
        shape = (c.height, c.width, 4)

        def _cb(_user, frame, ptr):
            callback(int(frame), np.ctypeslib.as_array(ptr, shape=shape))

        cb_c = N.FRAME_CALLBACK(_cb)
    N.check(N.lib().sar_render_sequence(
        renderer._h, C.byref(c), ang.ctypes.data_as(N._f64p), n, jobs_per_thread, seed & (2**64 - 1),
        N.SAR_SEQ_SHARED_POINTS if shared_points else 0,
        out.ctypes.data_as(N._u16p) if out is not None else None,
        C.cast(cb_c, C.c_void_p) if cb_c is not None else None, None))
    return out


def render_sequence_encoded(renderer: ParallelRenderer, config, angles: Sequence[float], jobs_per_thread: int,
                            pixel_format: PixelFormat, container: Container = Container.Raw, seed: Optional[int] = None,
                            shared_points: bool = False, callback=None) -> Optional[np.ndarray]:
    """render_sequence with every frame converted on the device (main.rs:52-57) and delivered encoded
    (PAM / BMP / raw bytes) — what the reference's encoder side threads write (main.rs:508-511).
    Returns [n_frames, n_bytes] uint8, or streams callback(frame_index, bytes_view) when given."""
    c = _pod(config)
    if seed is None:
        seed = int.from_bytes(os.urandom(8), "little")
    ang = np.ascontiguousarray(angles, dtype=np.float64)
    n = int(ang.shape[0])
    collected = None
    if container is Container.PngDeflate:          # frames of different sizes: callback only; without one, collect copies
        out = None
        if callback is None:
            collected = [None] * n

            def callback(frame, data, _dst=collected):
                _dst[frame] = data.copy()
    else:
        nb = N.lib().sar_encoded_size(c.width, c.height, pixel_format.value, container.value)
        if nb == 0:
            raise SarError(N.SAR_ERR_UNSUPPORTED, f"{pixel_format.name} cannot be written as {container.name}")
        out = np.empty((n, nb), dtype=np.uint8) if callback is None else None
    cb_c = None
    if callback is not None:
        def _cb(_user, frame, ptr, nbytes):
            callback(int(frame), np.ctypeslib.as_array(ptr, shape=(nbytes,)))

        cb_c = N.FRAME_BYTES_CALLBACK(_cb)
    N.check(N.lib().sar_render_sequence_encoded(
        renderer._h, C.byref(c), ang.ctypes.data_as(N._f64p), n, jobs_per_thread, seed & (2**64 - 1),
        N.SAR_SEQ_SHARED_POINTS if shared_points else 0, pixel_format.value, container.value,
        out.ctypes.data_as(N._u8p) if out is not None else None,
        C.cast(cb_c, C.c_void_p) if cb_c is not None else None, None))
    return collected if collected is not None else out
