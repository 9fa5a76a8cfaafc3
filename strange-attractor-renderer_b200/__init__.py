"""strange-attractor-renderer_b200 — B200-native replacement for the render path of
Icelk/strange-attractor-renderer (iterate → project → scatter → tone-map/colourise).

The directory name carries a hyphen (it mirrors the reference's crate name); import it as
`strange_attractor_renderer_b200` (a one-line alias package at the repo root).
"""
from .api import (  # noqa: F401
    BrighnessConstants, Colors, Config, EulerAxisRotation, FinalImage, Palette, ParallelRenderer,
    RenderKind, Runtime, SarConfig, SarError, Vec3, View, attractors, color_transforms, colorize,
    render, render_parallel, seed_points, angle_iter, angle_iter_files, render_sequence,
    PixelFormat, Container, encode_image, encode_png, write_image, render_sequence_encoded, AutoFrame, autoframe,
)
from . import _native, build  # noqa: F401

__all__ = [
    "BrighnessConstants", "Colors", "Config", "EulerAxisRotation", "FinalImage", "Palette",
    "ParallelRenderer", "RenderKind", "Runtime", "SarConfig", "SarError", "Vec3", "View",
    "attractors", "color_transforms", "colorize", "render", "render_parallel", "seed_points",
    "angle_iter", "angle_iter_files", "render_sequence", "PixelFormat", "Container", "encode_image", "encode_png", "write_image",
    "render_sequence_encoded", "AutoFrame", "autoframe",
]
