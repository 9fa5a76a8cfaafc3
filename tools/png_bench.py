"""Compressed PNG writer (sar_runtime_encode_png): size and time on a full-size frame.
    python tools/png_bench.py            # wall time of the blocking call, pinned output buffer
    ncu --metrics gpu__time_duration.sum --clock-control none --csv python tools/png_bench.py 1   # per-kernel times"""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import strange_attractor_renderer_b200 as S

N = S._native
L = N.lib()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
for (w, h, iters, preset) in ((1920, 1080, 1_000_000_000, "poisson_saturne"), (2048, 2048, 100_000_000, "solar_sail")):
    cfg = getattr(S.Config, preset)()
    cfg.width, cfg.height, cfg.iterations, cfg.transparent = w, h, iters, False
    if preset == "poisson_saturne":
        cfg.colors.brighness.offset = -0.25
    r = S.ParallelRenderer.new()
    S.render_parallel(r, cfg, 1, seed=9)
    rt = r.runtime()
    for fmt in (S.PixelFormat.Rgb16, S.PixelFormat.Rgb8):
        cap = L.sar_png_bound(w, h, fmt.value)
        host = C.c_void_p()
        N.check(L.sar_host_alloc(cap, C.byref(host)))
        out = C.cast(host, N._u8p)
        n = C.c_size_t()
        N.check(L.sar_runtime_encode_png(rt._h, fmt.value, out, cap, C.byref(n), None))
        t0 = time.perf_counter()
        for _ in range(reps):
            N.check(L.sar_runtime_encode_png(rt._h, fmt.value, out, cap, C.byref(n), None))
        dt = (time.perf_counter() - t0) / reps
        raw = h * (1 + w * (6 if fmt is S.PixelFormat.Rgb16 else 3))
        print(f"{preset} {w}x{h} {fmt.name}: scanlines {raw} -> file {n.value} bytes ({n.value / raw:.3f}), {dt * 1e3:.3f} ms per blocking call "
              f"({raw / dt / 1e9:.1f} GB/s of scanlines)")
        L.sar_host_free(host)
    r.shutdown()
