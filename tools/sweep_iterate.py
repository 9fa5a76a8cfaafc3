"""Sweep of the iterate kernel alone on one GPU: trajectories per thread (NT) x lanes per SM x image.

    python tools/sweep_iterate.py            # product library
    SWEEP_DIAG=1 python tools/sweep_iterate.py   # libsar_b200_diag.so: also mode 1 (arithmetic only) and 4 (no win path)

env: SWEEP_NT=1,2,4  SWEEP_PIPE=0,1  SWEEP_TILE=0|1 (shared-memory tile scatter for small images off / on)  SWEEP_LANES=768,896,1024  SWEEP_SHAPES=poisson:2048x2048,solar:1800x2000,poisson:4096x4096
     SWEEP_ITERS=1e9  SWEEP_MODES=0,1,4 (diag only)
Prints the median of 3 launches as G recorded iterations/s (warm-up steps excluded)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import strange_attractor_renderer_b200 as S
from strange_attractor_renderer_b200 import _native as N
from strange_attractor_renderer_b200 import build as B

DIAG = os.environ.get("SWEEP_DIAG", "0") == "1"
if DIAG:
    if not os.path.exists(B.OUT_DIAG):
        B.build(diagnostics=True)
    N.LIB_PATH = B.OUT_DIAG
L = N.lib()
stream = torch.cuda.Stream()
sp = C.c_void_p(stream.cuda_stream)
ITERS = int(float(os.environ.get("SWEEP_ITERS", "1e9")))
SMS = torch.cuda.get_device_properties(0).multi_processor_count
shapes = os.environ.get("SWEEP_SHAPES", "poisson:2048x2048,solar:1800x2000,poisson:4096x4096").split(",")
modes = [int(m) for m in os.environ.get("SWEEP_MODES", "0,1,4" if DIAG else "0").split(",")]
NAMES = {0: "product", 1: "arithmetic only", 2: "RED only", 4: "no win path", 5: "hot-pixel table cost model"}
if os.environ.get("SWEEP_TILE"):
    N.check(L.sar_set_option(b"tile_scatter", int(os.environ["SWEEP_TILE"])))
if DIAG and os.environ.get("SWEEP_HOT"):
    N.check(L.sar_set_option(b"diag_hot", int(os.environ["SWEEP_HOT"])))
for shape in shapes:
    preset, wh = shape.split(":")
    W, H = (int(v) for v in wh.split("x"))
    cfg = S.Config.poisson_saturne() if preset == "poisson" else S.Config.solar_sail()
    cfg.width, cfg.height = W, H
    if preset == "solar":
        cfg.angle = 3.839724354387525
    rt = C.c_void_p()
    N.check(L.sar_runtime_new(W, H, 0, C.byref(rt)))
    for mode in modes:
        if DIAG:
            N.check(L.sar_set_option(b"diagnostic_mode", mode))
        combos = [(nt, pipe) for nt in (int(v) for v in os.environ.get("SWEEP_NT", "1,2,4").split(","))
                  for pipe in (int(v) for v in os.environ.get("SWEEP_PIPE", "0,1").split(","))]
        for nt, pipe in combos:
            N.check(L.sar_set_option(b"traj_per_thread", nt))
            N.check(L.sar_set_option(b"pipeline", pipe))
            for lanes_per_sm in [int(v) for v in os.environ.get("SWEEP_LANES", "768,896,1024,1152").split(",")]:
                lanes = SMS * lanes_per_sm
                pod = cfg.to_pod()
                pod.iterations = ITERS // lanes
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
                print(f"{preset} {W}x{H} {NAMES.get(mode, mode)} NT {nt} pipe {pipe} lanes/SM {lanes_per_sm}: {ms:.3f} ms "
                      f"{pod.iterations * lanes / ms / 1e6:.2f} Git/s", flush=True)
    if DIAG:
        N.check(L.sar_set_option(b"diagnostic_mode", 0))
    L.sar_runtime_free(rt)
