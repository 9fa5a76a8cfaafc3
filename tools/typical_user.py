"""The reference CLI's default job (poisson-saturne, 1e7 iterations, 1920x1080, 12 jobs per thread,
main.rs:192-310) and a 1e8 variant through the public API, wall clock per frame, vs the CPU port."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import strange_attractor_renderer_b200 as S
from oracle import oracle as O

r = S.ParallelRenderer.new()
for iters in (10_000_000, 100_000_000, 1_000_000_000):
    cfg = S.Config.poisson_saturne(); cfg.iterations = iters
    out = np.empty((cfg.height, cfg.width, 4), np.uint16)
    S.render_parallel(r, cfg, 12, seed=1, out=out)
    t0 = time.perf_counter()
    for k in range(5):
        S.render_parallel(r, cfg, 12, seed=k, out=out)
    gpu = (time.perf_counter() - t0) / 5
    n, per_job = r.plan(iters, 12)
    threads = os.cpu_count() or 8
    ocfg = cfg.to_pod()
    t0 = time.perf_counter()
    O.render_parallel(ocfg, threads, 12, O.seed_points(1, 0, threads * 12))
    cpu = time.perf_counter() - t0
    print(f"{iters:.0e} iterations 1920x1080 jobs_per_thread 12: GPU {gpu*1e3:8.3f} ms/frame (num_threads {n}, {per_job} steps/job, pageable output buffer)   "
          f"CPU port {cpu*1e3:9.1f} ms ({threads} threads)   x{cpu/gpu:.0f}", flush=True)
r.shutdown()
