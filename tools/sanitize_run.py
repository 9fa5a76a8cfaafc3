"""Small tour of every kernel of the library, for compute-sanitizer (memcheck / initcheck / racecheck):
    compute-sanitizer --tool memcheck python tools/sanitize_run.py"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import strange_attractor_renderer_b200 as S

L = S._native.lib()
cfg = S.Config.solar_sail()
cfg.width, cfg.height, cfg.iterations, cfg.angle = 180, 150, 3_000, 220 * math.pi / 180
for nt, pipe in ((1, 0), (2, 1), (4, 0)):
    S._native.check(L.sar_set_option(b"traj_per_thread", nt))
    S._native.check(L.sar_set_option(b"pipeline", pipe))
    rt = S.Runtime.new(cfg)
    S.render(cfg, rt, initial_points=S.seed_points(3, 0, 333))
    img, f32 = S.colorize(cfg, rt, want_f32=True)
    count, steps, zbuf, mx = rt.download()
    rt2 = S.Runtime.new(cfg)
    rt2.upload(count, steps, zbuf)
    rt2.merge(rt)
    for fmt in S.PixelFormat:
        for cont in S.Container:
            try:
                S.encode_image(rt, fmt, cont)
            except S.SarError:
                pass
        S.encode_png(rt, fmt)
S._native.check(L.sar_set_option(b"traj_per_thread", 1))
S._native.check(L.sar_set_option(b"pipeline", 0))
# the shared-memory tile path: an image that fits a tile, more than one block of jobs, NaN trajectories included
tcfg = S.Config.solar_sail()
tcfg.width, tcfg.height, tcfg.iterations, tcfg.angle = 120, 100, 400, 1.0
trt = S.Runtime.new(tcfg)
S.render(tcfg, trt, initial_points=S.seed_points(8, 0, 2500))
S.colorize(tcfg, trt)
tile_hits = int(trt.download()[0].sum())
af = S.autoframe(cfg, n_jobs=500, iterations=500, seed=2)
r = S.ParallelRenderer.new(threads=128)
cfg.iterations = 600_000
frames = S.render_sequence(r, cfg, S.angle_iter(0.0, 40.0, 10.0), 2, seed=5, shared_points=True)
enc = S.render_sequence_encoded(r, cfg, S.angle_iter(0.0, 20.0, 10.0), 2, S.PixelFormat.Rgb8, S.Container.Bmp, seed=5)
one = S.render_parallel(r, cfg, 2, seed=5)
base = S.Config.poisson_saturne()
base.attractor = S.attractors.PolynomialSprott3Degree(base.attractor.x, base.attractor.y, base.attractor.z, [-0.05] + [0.0] * 9, [0.0] * 10, [0.0] * 9 + [-0.05])
base.color_transform = S.color_transforms.ScreenBlend([0.4, -0.3, 0.25, 1.5], offset=0.35, factor=0.9)
base.width, base.height, base.iterations = 100, 90, 2_000
rt = S.Runtime.new(base)
S.render(base, rt, initial_points=S.seed_points(1, 0, 100))
S.colorize(base, rt)
r.shutdown()
print("sanitize tour done", tile_hits, int(count.sum()), af.diverged, frames.shape, enc.shape, one.shape)
