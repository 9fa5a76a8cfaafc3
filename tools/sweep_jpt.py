"""Lanes x jobs-per-lane sweep of the iterate kernel (does overlapping one job's warm-up with other
warps' recorded phase pay?).  Recorded iterations/s, kernel alone."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import strange_attractor_renderer_b200 as S
from strange_attractor_renderer_b200 import _native as N

L = N.lib()
ITER, W, H = 1_000_000_000, 2048, 2048
cfg = S.Config.poisson_saturne(); cfg.width, cfg.height = W, H
stream = torch.cuda.Stream(); sp = C.c_void_p(stream.cuda_stream)
sms = torch.cuda.get_device_properties(0).multi_processor_count
rt = C.c_void_p(); N.check(L.sar_runtime_new(W, H, 0, C.byref(rt)))
N.check(L.sar_set_option(b"defer", int(os.environ.get("SAR_DEFER", "0"))))
for wps in (4, 5, 6, 7, 8):
    for jpt in (1, 2, 3):
        lanes = sms * 128 * wps
        jobs = lanes * jpt
        pod = cfg.to_pod(); pod.iterations = ITER // jobs
        ts = []
        for rep in range(3):
            N.check(L.sar_runtime_reset_async(rt, sp))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record(stream)
            N.check(L.sar_render_seeded_async(C.byref(pod), rt, 1234, 0, jobs, lanes, sp))
            e1.record(stream); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[1]
        print(f"warps/SMSP {wps} jobs/lane {jpt} jobs {jobs} iters/job {pod.iterations}: {ms:7.3f} ms {pod.iterations * jobs / ms / 1e6:7.2f} Git/s recorded", flush=True)
