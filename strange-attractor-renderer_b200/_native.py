"""ctypes binding of libsar_b200.so (include/sar.h).  No fallback: if the CUDA library is
missing or fails to load, importing the package's compute entry points raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsar_b200.so")

SAR_MAX_PALETTE = 16
SAR_IPC_HANDLE_BYTES = 64
SAR_OK, SAR_ERR_INVALID, SAR_ERR_DIMS, SAR_ERR_CUDA, SAR_ERR_NOMEM, SAR_ERR_UNSUPPORTED = 0, -1, -2, -3, -4, -5
SAR_RENDER_GAS, SAR_RENDER_DEPTH = 0, 1
SAR_CT_POISSON_SATURNE, SAR_CT_ADJUSTED_VELOCITY, SAR_CT_SCREEN_BLEND = 0, 1, 2
SAR_ATTRACTOR_SPROTT2, SAR_ATTRACTOR_SPROTT3 = 0, 1
SAR_SEQ_SHARED_POINTS = 1
SAR_PIX_RGBA16, SAR_PIX_RGB16, SAR_PIX_RGBA8, SAR_PIX_RGB8 = 0, 1, 2, 3
SAR_FILE_RAW, SAR_FILE_PAM, SAR_FILE_BMP, SAR_FILE_PNG, SAR_FILE_PNG_DEFLATE = 0, 1, 2, 3, 4
FRAME_BYTES_CALLBACK = C.CFUNCTYPE(None, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint8), C.c_size_t)
FRAME_CALLBACK = C.CFUNCTYPE(None, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint16))


class SarConfig(C.Structure):
    """`sar_config` of include/sar.h: POD form of Config<PolynomialSprott2Degree, _> (lib.rs:265-287)."""

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


class SarAutoframeResult(C.Structure):
    """`sar_autoframe_result` of include/sar.h."""

    _fields_ = [("box", C.c_double * 6), ("center_camera", C.c_double * 3), ("scale", C.c_double),
                ("diverged", C.c_uint64), ("n_jobs", C.c_uint64)]


class SarError(RuntimeError):
    """A non-zero sar_status.  (The reference panics at the same places: lib.rs:678, 709-710, 1024.)"""

    def __init__(self, code: int, msg: str):
        super().__init__(f"sar error {code}: {msg}")
        self.code = code


# every symbol include/sar.h declares: name -> (restype, argtypes)
_P = C.POINTER
_vp, _u8p, _u16p, _u32p, _f32p, _f64p = C.c_void_p, _P(C.c_uint8), _P(C.c_uint16), _P(C.c_uint32), _P(C.c_float), _P(C.c_double)
_cfgp = _P(SarConfig)
SYMBOLS = {
    "sar_abi_version": (C.c_uint32, []),
    "sar_last_error": (C.c_char_p, []),
    "sar_device_count": (C.c_int, [_P(C.c_int)]),
    "sar_set_option": (C.c_int, [C.c_char_p, C.c_int64]),
    "sar_default_threads": (C.c_int, [C.c_int, _u32p]),
    "sar_config_defaults": (C.c_int, [_cfgp]),
    "sar_config_poisson_saturne": (C.c_int, [_cfgp]),
    "sar_config_solar_sail": (C.c_int, [_cfgp]),
    "sar_seed_points": (C.c_int, [C.c_uint64, C.c_uint64, C.c_uint64, _f64p]),
    "sar_autoframe": (C.c_int, [_cfgp, C.c_int, C.c_uint64, _f64p, C.c_uint64, C.c_uint64, _P(SarAutoframeResult)]),
    "sar_runtime_new": (C.c_int, [C.c_uint32, C.c_uint32, C.c_int, _P(_vp)]),
    "sar_runtime_free": (None, [_vp]),
    "sar_runtime_reset": (C.c_int, [_vp]),
    "sar_runtime_merge": (C.c_int, [_vp, _vp]),
    "sar_runtime_dims": (C.c_int, [_vp, _u32p, _u32p, _P(C.c_int)]),
    "sar_runtime_download": (C.c_int, [_vp, _u32p, _f64p, _f32p, _u32p]),
    "sar_runtime_upload": (C.c_int, [_vp, _u32p, _f64p, _f32p]),
    "sar_render": (C.c_int, [_cfgp, _vp, _f64p, C.c_uint64]),
    "sar_render_seeded": (C.c_int, [_cfgp, _vp, C.c_uint64, C.c_uint64, C.c_uint64]),
    "sar_colorize": (C.c_int, [_cfgp, _vp, _u16p, _f32p]),
    "sar_renderer_new": (C.c_int, [_P(C.c_int), C.c_int, C.c_uint32, _P(_vp)]),
    "sar_renderer_shutdown": (None, [_vp]),
    "sar_renderer_num_threads": (C.c_int, [_vp, _P(C.c_uint64)]),
    "sar_renderer_num_threads_for": (C.c_int, [_vp, C.c_uint64, _P(C.c_uint64)]),
    "sar_renderer_plan": (C.c_int, [_vp, C.c_uint64, C.c_uint64, _P(C.c_uint64), _P(C.c_uint64)]),
    "sar_render_parallel": (C.c_int, [_vp, _cfgp, C.c_uint64, C.c_uint64, _f64p, _u16p]),
    "sar_renderer_runtime": (C.c_int, [_vp, _P(_vp)]),
    "sar_render_sequence": (C.c_int, [_vp, _cfgp, _f64p, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint32, _u16p, _vp, _vp]),
    "sar_encoded_size": (C.c_size_t, [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]),
    "sar_encode_header": (C.c_int, [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, _u8p, C.c_size_t, _P(C.c_size_t)]),
    "sar_runtime_encode": (C.c_int, [_vp, C.c_uint32, C.c_uint32, _u8p, C.c_size_t, _vp]),
    "sar_png_bound": (C.c_size_t, [C.c_uint32, C.c_uint32, C.c_uint32]),
    "sar_runtime_encode_png": (C.c_int, [_vp, C.c_uint32, _u8p, C.c_size_t, _P(C.c_size_t), _vp]),
    "sar_write_file": (C.c_int, [C.c_char_p, _u8p, C.c_size_t]),
    "sar_render_sequence_encoded": (C.c_int, [_vp, _cfgp, _f64p, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, _u8p, _vp, _vp]),
    "sar_render_seeded_async": (C.c_int, [_cfgp, _vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, _vp]),
    "sar_render_device_async": (C.c_int, [_cfgp, _vp, _vp, C.c_uint64, C.c_uint64, C.c_uint32, _vp]),
    "sar_runtime_reset_async": (C.c_int, [_vp, _vp]),
    "sar_runtime_max_async": (C.c_int, [_vp, C.c_uint32, C.c_uint32, _vp]),
    "sar_runtime_get_max": (C.c_int, [_vp, _u32p, _vp]),
    "sar_runtime_set_max": (C.c_int, [_vp, C.c_uint32, _vp]),
    "sar_colorize_rows_async": (C.c_int, [_cfgp, _vp, C.c_uint32, C.c_uint32, _vp, _vp]),
    "sar_runtime_image_download": (C.c_int, [_vp, C.c_uint32, C.c_uint32, _vp, _vp]),
    "sar_stream_synchronize": (C.c_int, [_vp, _vp]),
    "sar_runtime_get_job_base": (C.c_int, [_vp, _P(C.c_uint64)]),
    "sar_runtime_set_job_base": (C.c_int, [_vp, C.c_uint64]),
    "sar_launch_count": (C.c_uint64, []),
    "sar_host_alloc": (C.c_int, [C.c_size_t, _P(_vp)]),
    "sar_host_free": (None, [_vp]),
    "sar_runtime_ipc_export": (C.c_int, [_vp, _u8p]),
    "sar_peer_open": (C.c_int, [_u8p, C.c_uint32, C.c_uint32, C.c_int, _P(_vp)]),
    "sar_peer_close": (None, [_vp]),
    "sar_runtime_merge_peers_async": (C.c_int, [_vp, _P(_vp), C.c_int, C.c_uint32, C.c_uint32, _vp]),
    "sar_frame_reset_async": (C.c_int, [_vp, C.c_int, C.c_uint32, _vp]),
    "sar_frame_export_async": (C.c_int, [_vp, _P(_vp), C.c_int, C.c_int, C.c_uint32, _vp]),
    "sar_frame_merge_async": (C.c_int, [_vp, _P(_vp), C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, _vp]),
    "sar_frame_colorize_async": (C.c_int, [_cfgp, _vp, _P(_vp), C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, _vp]),
    "sar_frame_image_wait_async": (C.c_int, [_vp, C.c_int, C.c_uint32, _vp]),
    "sar_frame_image_release_async": (C.c_int, [_vp, _P(_vp), C.c_int, C.c_int, C.c_uint32, _vp]),
    "sar_runtime_sync_error": (C.c_int, [_vp, _u32p, C.c_int]),
}

_lib = None


def lib() -> C.CDLL:
    """Load libsar_b200.so (built by build.py / __graft_entry__.build()).  Raises if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "There is no CPU fallback for the render path.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError = ABI mismatch, loud by design
            fn.restype, fn.argtypes = res, args
        if L.sar_abi_version() != 2:
            raise RuntimeError(f"libsar_b200.so ABI {L.sar_abi_version()} != 2")
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise SarError(rc, lib().sar_last_error().decode("utf-8", "replace"))
