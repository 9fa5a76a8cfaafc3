"""A few launches of the iterate kernel (+ max + colorize) for ncu.
Usage: prof_one.py [lanes_per_sm] [traj_per_thread] [preset] [WxH]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import strange_attractor_renderer_b200 as S
from strange_attractor_renderer_b200 import _native as N

L = N.lib()
lanes_per_sm = int(sys.argv[1]) if len(sys.argv) > 1 else 0
nt = int(sys.argv[2]) if len(sys.argv) > 2 else 0
preset = sys.argv[3] if len(sys.argv) > 3 else "poisson"
W, H = (int(v) for v in sys.argv[4].split("x")) if len(sys.argv) > 4 else (2048, 2048)
cfg = S.Config.poisson_saturne() if preset == "poisson" else S.Config.solar_sail()
cfg.width, cfg.height = W, H
sms = torch.cuda.get_device_properties(0).multi_processor_count
t = C.c_uint32()
N.check(L.sar_default_threads(0, C.byref(t)))
lanes = sms * lanes_per_sm if lanes_per_sm else int(t.value)
pod = cfg.to_pod()
pod.iterations = 1_000_000_000 // lanes
if nt:
    N.check(L.sar_set_option(b"traj_per_thread", nt))
if os.environ.get("SAR_PIPE"):
    N.check(L.sar_set_option(b"pipeline", int(os.environ["SAR_PIPE"])))
rt = C.c_void_p()
N.check(L.sar_runtime_new(W, H, 0, C.byref(rt)))
for _ in range(3):
    N.check(L.sar_runtime_reset_async(rt, None))
    N.check(L.sar_render_seeded_async(C.byref(pod), rt, 1234, 0, lanes, lanes, None))
    N.check(L.sar_runtime_max_async(rt, 0, 0, None))
    N.check(L.sar_colorize_rows_async(C.byref(pod), rt, 0, 0, None, None))
    N.check(L.sar_stream_synchronize(rt, None))
print("done", lanes, pod.iterations)
