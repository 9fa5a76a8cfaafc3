"""lanes/SM sweep of the iterate kernel alone for three image shapes (product path and, as a
ceiling, diagnostic mode 4 = same atomic but no win path).  Usage: python tools/sweep_lanes.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import strange_attractor_renderer_b200 as S
from strange_attractor_renderer_b200 import _native as N
L = N.lib()
stream = torch.cuda.Stream(); sp = C.c_void_p(stream.cuda_stream)
ITERS = int(float(os.environ.get("SWEEP_ITERS", "1e9")))
for preset, W, H in (("poisson", 2048, 2048), ("solar", 1800, 2000), ("poisson", 4096, 4096)):
    cfg = S.Config.poisson_saturne() if preset == "poisson" else S.Config.solar_sail()
    cfg.width, cfg.height = W, H
    if preset == "solar": cfg.angle = 3.839724354387525
    rt = C.c_void_p(); N.check(L.sar_runtime_new(W, H, 0, C.byref(rt)))
    for mode in [int(m) for m in os.environ.get("SWEEP_MODES", "0,4").split(",")]:
        N.check(L.sar_set_option(b"diagnostic_mode", mode))
        for lanes_per_sm in [int(v) for v in os.environ.get("SWEEP_LANES", "640,768,896,1024,1152").split(",")]:
            lanes = 148 * lanes_per_sm
            pod = cfg.to_pod(); pod.iterations = ITERS // lanes
            ts = []
            for rep in range(3):
                N.check(L.sar_runtime_reset_async(rt, sp))
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(); e0.record(stream)
                N.check(L.sar_render_seeded_async(C.byref(pod), rt, 1234, 0, lanes, lanes, sp))
                e1.record(stream); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
            ms = sorted(ts)[1]
            print(f"{preset} {W}x{H} {'product' if mode == 0 else 'no win path'} lanes/SM {lanes_per_sm}: {ms:.3f} ms {pod.iterations*lanes/ms/1e6:.2f} Git/s", flush=True)
    N.check(L.sar_set_option(b"diagnostic_mode", 0))
    L.sar_runtime_free(rt)
