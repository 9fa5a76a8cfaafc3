"""Tuning sweep for the iterate kernel on one GPU: deferred-test depth x lanes.
Prints recorded iterations/s of the iterate kernel alone (CUDA events), and checks that every
variant produces the same buffers as depth 0 on a small job list (results must not depend on it).
Usage: python tools/sweep_iterate.py [preset] [iterations] [W] [H]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import strange_attractor_renderer_b200 as S
from strange_attractor_renderer_b200 import _native as N

L = N.lib()
preset = sys.argv[1] if len(sys.argv) > 1 else "poisson"
ITER = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1_000_000_000
W = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
H = int(sys.argv[4]) if len(sys.argv) > 4 else 2048
cfg = S.Config.poisson_saturne() if preset == "poisson" else S.Config.solar_sail()
if preset != "poisson":
    cfg.angle = 3.839724354387525
cfg.width, cfg.height = W, H
stream = torch.cuda.Stream()
sp = C.c_void_p(stream.cuda_stream)
sms = torch.cuda.get_device_properties(0).multi_processor_count

# parity across depths on a small problem
small = S.Config.solar_sail(); small.width, small.height, small.iterations = 300, 300, 5001
ref = None
for d in range(5):
    N.check(L.sar_set_option(b"defer", d))
    rt = S.Runtime.new(small)
    S.render(small, rt, initial_points=S.seed_points(3, 0, 700))
    st = rt.download()
    if ref is None:
        ref = st
    else:
        same = all(np.array_equal(a.view(np.uint8), b.view(np.uint8)) for a, b in zip(ref[:3], st[:3])) and ref[3] == st[3]
        print(f"defer {d} == defer 0: {same}")
    rt.close()

rt = C.c_void_p()
N.check(L.sar_runtime_new(W, H, 0, C.byref(rt)))
for d in [int(x) for x in os.environ.get("SWEEP_DEFER", "0,1,2,3,4").split(",") if x != ""]:
    N.check(L.sar_set_option(b"defer", d))
    for wps in (2, 3, 4, 6, 8):
        lanes = sms * 4 * 32 * wps
        pod = cfg.to_pod()
        pod.iterations = ITER // lanes
        rec = pod.iterations * lanes
        ts = []
        for rep in range(3):
            N.check(L.sar_runtime_reset_async(rt, sp))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record(stream)
            N.check(L.sar_render_seeded_async(C.byref(pod), rt, 1234, 0, lanes, lanes, sp))
            e1.record(stream)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[1]
        print(f"defer {d} warps/SMSP {wps} lanes {lanes} iters/job {pod.iterations}: {ms:8.3f} ms  {rec / ms / 1e6:8.2f} Git/s (recorded)  "
              f"{(rec + 1000 * lanes) / ms / 1e6:8.2f} Git/s (incl. warm-up)", flush=True)

print("--- diagnostic modes at defer 1 (1 = arithmetic only, 2 = RED only, 3 = RED + 4-byte hint load) ---")
for mode in (1, 2, 3, 4, 5, 6, 0):
    N.check(L.sar_set_option(b"diagnostic_mode", mode))
    N.check(L.sar_set_option(b"defer", 1))
    for wps in (2, 4, 8):
        lanes = sms * 4 * 32 * wps
        pod = cfg.to_pod()
        pod.iterations = ITER // lanes
        rec = pod.iterations * lanes
        ts = []
        for rep in range(3):
            N.check(L.sar_runtime_reset_async(rt, sp))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record(stream)
            N.check(L.sar_render_seeded_async(C.byref(pod), rt, 1234, 0, lanes, lanes, sp))
            e1.record(stream)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[1]
        print(f"mode {mode} warps/SMSP {wps}: {ms:8.3f} ms  {rec / ms / 1e6:8.2f} Git/s (recorded)  "
              f"{(rec + 1000 * lanes) / ms / 1e6:8.2f} Git/s (incl. warm-up)", flush=True)
N.check(L.sar_set_option(b"diagnostic_mode", 0))
