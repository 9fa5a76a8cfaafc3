import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np
import strange_attractor_renderer_b200 as S
from oracle import oracle as O
L = S._native.lib()
S._native.check(L.sar_set_option(b"tile_scatter", 2))
for (w, h, iters, jobs, preset) in ((256, 256, 300, 20000, "poisson"), (512, 512, 200, 30000, "solar"), (640, 640, 100, 16000, "poisson"), (300, 333, 150, 15000, "solar")):
    cfg = S.Config.poisson_saturne() if preset == "poisson" else S.Config.solar_sail()
    cfg.width, cfg.height, cfg.iterations, cfg.angle = w, h, iters, 0.7
    pts = S.seed_points(5, 0, jobs)
    rt = S.Runtime.new(cfg)
    S.render(cfg, rt, initial_points=pts)
    count, steps, zbuf, mx = rt.download()
    ort = O.Runtime(w, h)
    O.render_jobs_mt(cfg.to_pod(), ort, pts)
    ok = (np.array_equal(count, ort.count), np.array_equal(zbuf.view(np.uint32), ort.zbuf.view(np.uint32)),
          np.array_equal(steps.view(np.uint64), ort.steps.view(np.uint64)), mx == ort.max,
          np.array_equal(S.colorize(cfg, rt), O.colorize(cfg.to_pod(), ort)))
    print(preset, w, h, "count/zbuf/steps/max/image:", ok, "recorded", int(count.sum()), flush=True)
