"""ctypes loader for the CPU oracle (oracle/sar_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference legs, and the
N-rank parity frames it checks outside every timed region) may import this module.  The product
package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")

SAR_MAX_PALETTE = 16


class SarConfig(C.Structure):
    """Mirror of `sar_config` (include/sar.h) — itself the POD form of Config, lib.rs:265-287."""

    _fields_ = [
        ("iterations", C.c_uint64),
        ("width", C.c_uint32),
        ("height", C.c_uint32),
        ("render_kind", C.c_uint32),
        ("transparent", C.c_uint32),
        ("silent", C.c_uint32),
        ("ct_kind", C.c_uint32),
        ("angle", C.c_double),
        ("coef", (C.c_double * 10) * 3),
        ("center_camera", C.c_double * 3),
        ("axis", C.c_double * 3),
        ("rotation", C.c_double),
        ("scale", C.c_double),
        ("ct_offset", C.c_double),
        ("ct_factor", C.c_double),
        ("palette_len", C.c_uint32),
        ("attractor_kind", C.c_uint32),
        ("palette_rgb", (C.c_double * 3) * SAR_MAX_PALETTE),
        ("bright_offset", C.c_double),
        ("bright_factor", C.c_double),
        ("coef3", (C.c_double * 10) * 3),
        ("ct_weights", C.c_double * 4),
    ]

    def copy(self) -> "SarConfig":
        return SarConfig.from_buffer_copy(bytes(self))


class _OrcRuntime(C.Structure):
    _fields_ = [
        ("w", C.c_uint32),
        ("h", C.c_uint32),
        ("count", C.POINTER(C.c_uint32)),
        ("steps", C.POINTER(C.c_double)),
        ("zbuf", C.POINTER(C.c_float)),
        ("max", C.c_uint32),
    ]


class OrcStats(C.Structure):
    _fields_ = [
        ("recorded", C.c_uint64),
        ("z_wins", C.c_uint64),
        ("nan_iters", C.c_uint64),
        ("z_ties", C.c_uint64),
    ]


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (gcc, -ffp-contract=off)."""
    src = os.path.join(_HERE, "sar_oracle.c")
    stale = (not os.path.exists(_SO)) or any(
        os.path.getmtime(p) > os.path.getmtime(_SO)
        for p in (src, os.path.join(_HERE, "sar_oracle.h"), os.path.join(_HERE, "..", "include", "sar.h"))
    )
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True, capture_output=True)
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        P = C.POINTER
        dp = P(C.c_double)
        L.orc_next_point.argtypes = [P((C.c_double * 10) * 3), dp]
        L.orc_next_point_cfg.argtypes = [P(SarConfig), dp]
        L.orc_next_point_cfg.restype = None
        L.orc_rotation_matrix.argtypes = [dp, C.c_double, dp]
        L.orc_mul_right.argtypes = [dp, dp, dp]
        L.orc_color_transform.argtypes = [P(SarConfig), dp, dp]
        L.orc_color_transform.restype = C.c_double
        L.orc_palette_interpolate.argtypes = [P(SarConfig), C.c_double, dp]
        L.orc_runtime_new.argtypes = [C.c_uint32, C.c_uint32, P(P(_OrcRuntime))]
        L.orc_runtime_free.argtypes = [P(_OrcRuntime)]
        L.orc_runtime_reset.argtypes = [P(_OrcRuntime)]
        L.orc_runtime_merge.argtypes = [P(_OrcRuntime), P(_OrcRuntime)]
        L.orc_render.argtypes = [P(SarConfig), P(_OrcRuntime), dp, P(OrcStats)]
        L.orc_render_jobs.argtypes = [P(SarConfig), P(_OrcRuntime), dp, C.c_uint64, P(OrcStats)]
        L.orc_render_jobs_mt.argtypes = [P(SarConfig), P(_OrcRuntime), dp, C.c_uint64, C.c_uint32, P(OrcStats)]
        L.orc_render_jobs_mt.restype = C.c_int
        L.orc_colorize.argtypes = [P(SarConfig), P(_OrcRuntime), P(C.c_uint16), dp]
        L.orc_render_parallel.argtypes = [P(SarConfig), C.c_uint32, C.c_uint64, dp, P(C.c_uint16), P(P(_OrcRuntime))]
        L.orc_encode.argtypes = [P(C.c_uint16), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, P(C.c_uint8)]
        L.orc_encode.restype = C.c_size_t
        L.orc_screen_bbox.argtypes = [P(SarConfig), dp, C.c_uint64, dp]
        L.orc_screen_bbox_jobs.argtypes = [P(SarConfig), dp, C.c_uint64, C.c_uint64, dp, P(C.c_uint64)]
        L.orc_screen_bbox_jobs.restype = None
        L.orc_seed_points.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, dp]
        L.orc_config_poisson_saturne.argtypes = [P(SarConfig)]
        L.orc_config_solar_sail.argtypes = [P(SarConfig)]
        for f in ("orc_next_point", "orc_rotation_matrix", "orc_mul_right", "orc_palette_interpolate",
                  "orc_runtime_free", "orc_runtime_reset", "orc_render", "orc_render_jobs", "orc_colorize",
                  "orc_screen_bbox", "orc_seed_points", "orc_config_poisson_saturne", "orc_config_solar_sail"):
            getattr(L, f).restype = None
        _lib = L
    return _lib


def _dp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def poisson_saturne() -> SarConfig:
    c = SarConfig()
    lib().orc_config_poisson_saturne(C.byref(c))
    return c


def solar_sail() -> SarConfig:
    c = SarConfig()
    lib().orc_config_solar_sail(C.byref(c))
    return c


def as_oracle_config(cfg) -> SarConfig:
    """Accept an oracle SarConfig or any ctypes struct with the same layout (the package's)."""
    if isinstance(cfg, SarConfig):
        return cfg
    raw = bytes(cfg)
    assert len(raw) == C.sizeof(SarConfig), (len(raw), C.sizeof(SarConfig))
    return SarConfig.from_buffer_copy(raw)


def seed_points(seed: int, first: int, n: int) -> np.ndarray:
    out = np.empty((n, 3), dtype=np.float64)
    lib().orc_seed_points(seed, first, n, _dp(out))
    return out


def next_point(cfg: SarConfig, p) -> np.ndarray:
    q = np.array(p, dtype=np.float64)
    lib().orc_next_point_cfg(C.byref(cfg), _dp(q))
    return q


def rotation_matrix(axis, rotation: float) -> np.ndarray:
    a = np.array(axis, dtype=np.float64)
    m = np.empty((3, 3), dtype=np.float64)
    lib().orc_rotation_matrix(_dp(a), rotation, _dp(m))
    return m


def color_transform(cfg: SarConfig, delta, screen) -> float:
    d = np.array(delta, dtype=np.float64)
    s = np.array(screen, dtype=np.float64)
    return float(lib().orc_color_transform(C.byref(cfg), _dp(d), _dp(s)))


def palette_interpolate(cfg: SarConfig, value: float) -> np.ndarray:
    rgb = np.empty(3, dtype=np.float64)
    lib().orc_palette_interpolate(C.byref(cfg), value, _dp(rgb))
    return rgb


def screen_bbox(cfg: SarConfig, init, n: int) -> np.ndarray:
    box = np.empty(6, dtype=np.float64)
    p = np.array(init, dtype=np.float64)
    lib().orc_screen_bbox(C.byref(cfg), _dp(p), n, _dp(box))
    return box


class Runtime:
    """Owning wrapper of orc_runtime; numpy views of the three textures."""

    def __init__(self, w: int, h: int, _ptr=None):
        if _ptr is None:
            p = C.POINTER(_OrcRuntime)()
            if lib().orc_runtime_new(w, h, C.byref(p)) != 0:
                raise MemoryError("orc_runtime_new")
            _ptr = p
        self._p = _ptr
        self.w, self.h = int(self._p.contents.w), int(self._p.contents.h)

    def __del__(self):
        if getattr(self, "_p", None):
            lib().orc_runtime_free(self._p)
            self._p = None

    @property
    def count(self) -> np.ndarray:
        return np.ctypeslib.as_array(self._p.contents.count, shape=(self.h, self.w))

    @property
    def steps(self) -> np.ndarray:
        return np.ctypeslib.as_array(self._p.contents.steps, shape=(self.h, self.w))

    @property
    def zbuf(self) -> np.ndarray:
        return np.ctypeslib.as_array(self._p.contents.zbuf, shape=(self.h, self.w))

    @property
    def max(self) -> int:
        return int(self._p.contents.max)

    def reset(self) -> None:
        lib().orc_runtime_reset(self._p)

    def merge(self, other: "Runtime") -> None:
        if lib().orc_runtime_merge(self._p, other._p) != 0:
            raise ValueError("dimension mismatch")  # reference: assert_eq! panic, lib.rs:709-710

    def load(self, count, steps, zbuf) -> None:
        """Overwrite the textures (used to re-create a state downloaded from the GPU)."""
        self.count[...] = count
        self.steps[...] = steps
        self.zbuf[...] = zbuf
        self._p.contents.max = int(self.count.max()) if self.count.size else 0


def render(cfg, rt: Runtime, init, stats: OrcStats | None = None) -> None:
    """One reference render() call (lib.rs:747) from start point `init` (3 f64)."""
    cfg = as_oracle_config(cfg)
    p = np.ascontiguousarray(init, dtype=np.float64).reshape(3)
    lib().orc_render(C.byref(cfg), rt._p, _dp(p), C.byref(stats) if stats is not None else None)


def render_jobs(cfg, rt: Runtime, init_xyz, stats: OrcStats | None = None) -> None:
    cfg = as_oracle_config(cfg)
    pts = np.ascontiguousarray(init_xyz, dtype=np.float64).reshape(-1, 3)
    lib().orc_render_jobs(C.byref(cfg), rt._p, _dp(pts), pts.shape[0],
                          C.byref(stats) if stats is not None else None)


def render_jobs_mt(cfg, rt: Runtime, init_xyz, n_threads: int = 0, stats: OrcStats | None = None) -> None:
    """render_jobs on n_threads OS threads (0 = all cores); bit-identical to render_jobs (sar_oracle.c)."""
    cfg = as_oracle_config(cfg)
    pts = np.ascontiguousarray(init_xyz, dtype=np.float64).reshape(-1, 3)
    if n_threads <= 0:
        n_threads = os.cpu_count() or 8
    if lib().orc_render_jobs_mt(C.byref(cfg), rt._p, _dp(pts), pts.shape[0], n_threads,
                                C.byref(stats) if stats is not None else None) != 0:
        raise RuntimeError("orc_render_jobs_mt failed")


def colorize(cfg, rt: Runtime, want_f64: bool = False):
    cfg = as_oracle_config(cfg)
    out = np.empty((rt.h, rt.w, 4), dtype=np.uint16)
    f = np.empty((rt.h, rt.w, 4), dtype=np.float64) if want_f64 else None
    lib().orc_colorize(C.byref(cfg), rt._p, out.ctypes.data_as(C.POINTER(C.c_uint16)),
                       _dp(f) if f is not None else None)
    return (out, f) if want_f64 else out


def render_parallel(cfg, n_threads: int, jobs_per_thread: int, init_xyz, want_runtime: bool = False):
    """render_parallel (lib.rs:1051) on n_threads OS threads; returns rgba (and the merged Runtime)."""
    cfg = as_oracle_config(cfg)
    pts = np.ascontiguousarray(init_xyz, dtype=np.float64).reshape(-1, 3)
    assert pts.shape[0] >= n_threads * jobs_per_thread
    out = np.empty((cfg.height, cfg.width, 4), dtype=np.uint16)
    mp = C.POINTER(_OrcRuntime)()
    rc = lib().orc_render_parallel(C.byref(cfg), n_threads, jobs_per_thread, _dp(pts),
                                   out.ctypes.data_as(C.POINTER(C.c_uint16)),
                                   C.byref(mp) if want_runtime else None)
    if rc != 0:
        raise RuntimeError("orc_render_parallel failed")
    return (out, Runtime(0, 0, _ptr=mp)) if want_runtime else out


def encode(rgba_u16: np.ndarray, fmt: int, container: int):
    """main.rs:52-98 on the host: converted + containerised bytes of an RGBA16 image (None if unsupported)."""
    img = np.ascontiguousarray(rgba_u16, dtype=np.uint16)
    h, w = img.shape[:2]
    n = lib().orc_encode(img.ctypes.data_as(C.POINTER(C.c_uint16)), w, h, fmt, container, None)
    if n == 0:
        return None
    out = np.empty(n, dtype=np.uint8)
    lib().orc_encode(img.ctypes.data_as(C.POINTER(C.c_uint16)), w, h, fmt, container, out.ctypes.data_as(C.POINTER(C.c_uint8)))
    return out


def screen_bbox_jobs(cfg, init_xyz, n: int):
    """-> (box[6], diverged): union of the screen-space boxes of the bounded trajectories."""
    cfg = as_oracle_config(cfg)
    pts = np.ascontiguousarray(init_xyz, dtype=np.float64).reshape(-1, 3)
    box = np.empty(6, dtype=np.float64)
    bad = C.c_uint64()
    lib().orc_screen_bbox_jobs(C.byref(cfg), _dp(pts), pts.shape[0], n, _dp(box), C.byref(bad))
    return box, int(bad.value)
