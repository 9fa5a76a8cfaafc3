"""BASELINE configs[0] (the reference's own CPU-runnable case): poisson-saturne, 1e7 iterations, 512x512, render() = ONE
serial trajectory (lib.rs:747-838).  Times the oracle on one host thread and the same call on the GPU (one lane — exact,
and by construction not what a GPU is for; render_parallel is the fast path) and checks they agree bit for bit."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import strange_attractor_renderer_b200 as S
from oracle import oracle as O          # checker / CPU baseline only

cfg = S.Config.poisson_saturne()
cfg.width, cfg.height, cfg.iterations = 512, 512, 10_000_000
pts = S.seed_points(2024, 0, 1)
ort = O.Runtime(512, 512)
t0 = time.perf_counter()
O.render_jobs(cfg.to_pod(), ort, pts)
oimg = O.colorize(cfg.to_pod(), ort)
t_cpu = time.perf_counter() - t0
rt = S.Runtime.new(cfg)
S.render(cfg, rt, initial_points=pts)    # warm-up of the context / tables
rt.reset()
t0 = time.perf_counter()
S.render(cfg, rt, initial_points=pts)
img = S.colorize(cfg, rt)
t_gpu = time.perf_counter() - t0
r = S.ParallelRenderer.new()
S.render_parallel(r, cfg, 1, seed=1)
t0 = time.perf_counter()
S.render_parallel(r, cfg, 1, seed=1)
t_par = time.perf_counter() - t0
n, per_job = r.plan(cfg.iterations, 1)
print(f"cfg0 1e7 @ 512x512: CPU oracle, 1 thread: {t_cpu * 1e3:.1f} ms ({1e7 / t_cpu:.3g} it/s); GPU render() one lane: {t_gpu * 1e3:.1f} ms "
      f"({1e7 / t_gpu:.3g} it/s), image equal: {bool(np.array_equal(img, oimg))}; GPU render_parallel ({n} jobs x {per_job}): {t_par * 1e3:.2f} ms "
      f"({n * per_job / t_par:.3g} it/s)")
