"""Measure BASELINE.json's other configurations on this box's GPU(s) (not bench lines; evidence for
profiles/).  cfg 2 (solar-sail 1e9 1800x2000 220 deg) as device-resident frames, cfg 4 shape
(poisson 4096x4096) per GPU, cfg 5 (solar-sail 360-frame sweep, 1e8/frame, 2048x2048) through
sar_render_sequence.  Usage: python tools/run_configs.py [frames_for_cfg5]"""
import ctypes as C
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import strange_attractor_renderer_b200 as S
from strange_attractor_renderer_b200 import _native as N
from strange_attractor_renderer_b200 import dist as D

L = N.lib()
stream = torch.cuda.Stream()
sp = C.c_void_p(stream.cuda_stream)


def frames(cfg, iters, reps=5, lanes=0):
    fr = D.Frame(cfg, device=0, world=1, rank=0, group=None, lanes=lanes, jobs_per_thread=1, iterations_per_gpu=iters, seed=1234)
    for _ in range(2):
        fr.step_device(sp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fr.step_device(sp)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    count = np.empty((fr.h, fr.w), np.uint32)
    N.check(L.sar_runtime_download(fr.rt, count.ctypes.data_as(N._u32p), None, None, None))
    total = fr.recorded_iterations_total()
    nan = int(count[0, 0])
    lit = float((count > 0).mean())
    fr.close()
    return ms, total, nan, lit, fr.lanes, fr.iterations_per_job


print("| config | lanes | iters/job | ms/frame | iterations/s | count[(0,0)] share | lit pixels |")
print("|---|---|---|---|---|---|---|")
cfg = S.Config.poisson_saturne(); cfg.width = cfg.height = 2048
ms, tot, nan, lit, lanes, ipj = frames(cfg, 10**9)
print(f"| cfg1 poisson-saturne 1e9 2048x2048 | {lanes} | {ipj} | {ms:.3f} | {tot / ms * 1e3:.4g} | {nan / tot:.4f} | {lit:.4f} |", flush=True)
cfg = S.Config.solar_sail(); cfg.width, cfg.height, cfg.angle = 1800, 2000, 220 * math.pi / 180
ms, tot, nan, lit, lanes, ipj = frames(cfg, 10**9)
print(f"| cfg2 solar-sail 1e9 1800x2000 220deg | {lanes} | {ipj} | {ms:.3f} | {tot / ms * 1e3:.4g} | {nan / tot:.4f} | {lit:.4f} |", flush=True)
cfg = S.Config.poisson_saturne(); cfg.width = cfg.height = 4096
ms, tot, nan, lit, lanes, ipj = frames(cfg, 10**9, reps=3)
print(f"| cfg3 shape: poisson-saturne 1e9 per GPU, 4096x4096 | {lanes} | {ipj} | {ms:.3f} | {tot / ms * 1e3:.4g} | {nan / tot:.4f} | {lit:.4f} |", flush=True)

# cfg 4: the sweep
nfr = int(sys.argv[1]) if len(sys.argv) > 1 else 36
cfg = S.Config.solar_sail(); cfg.width = cfg.height = 2048; cfg.iterations = 100_000_000
angles = S.angle_iter(0.0, 360.0, 360.0 / nfr)
print()
print("| cfg4 sweep: solar-sail 1e8/frame 2048x2048 | threads | points | frames | s total | ms/frame | iterations/s |")
print("|---|---|---|---|---|---|---|")
for threads in (0, 75776, 37888, 18944):
    for shared in (False, True):
        r = S.ParallelRenderer.new(threads=threads)
        n = r.num_threads()
        S.render_sequence(r, cfg, angles[:2], 1, seed=7, shared_points=shared, callback=lambda f, im: None)
        t0 = time.perf_counter()
        got = []
        S.render_sequence(r, cfg, angles, 1, seed=7, shared_points=shared, callback=lambda f, im: got.append(int(im[0, 0, 0])))
        dt = time.perf_counter() - t0
        rec = (100_000_000 // n) * n * len(angles)
        print(f"| | {n} | {'shared (warm-up once)' if shared else 'fresh per frame'} | {len(got)} | {dt:.3f} | {1e3 * dt / len(angles):.3f} | {rec / dt:.4g} |", flush=True)
        r.shutdown()
