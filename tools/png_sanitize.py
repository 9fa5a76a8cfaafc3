"""The compressed-PNG kernels alone, for compute-sanitizer (racecheck is slow on the full tour):
    compute-sanitizer --tool racecheck python tools/png_sanitize.py"""
import os
import sys
import zlib

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import strange_attractor_renderer_b200 as S

cfg = S.Config.solar_sail()
cfg.width, cfg.height, cfg.iterations, cfg.transparent = 640, 300, 2_000, False
rt = S.Runtime.new(cfg)
S.render(cfg, rt, initial_points=S.seed_points(3, 0, 2000))
S.colorize(cfg, rt)
for fmt in S.PixelFormat:
    png = S.encode_png(rt, fmt).tobytes()
    n = int.from_bytes(png[33:37], "big")
    assert zlib.crc32(png[37:41 + n]) == int.from_bytes(png[41 + n:45 + n], "big")
    zlib.decompress(png[41:41 + n])
print("png sanitize tour done")
