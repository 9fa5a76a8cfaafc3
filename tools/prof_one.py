"""A few launches of the iterate kernel for ncu.  Usage: prof_one.py [warps_per_smsp] [defer] [preset]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import strange_attractor_renderer_b200 as S
from strange_attractor_renderer_b200 import _native as N

L = N.lib()
wps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
defer = int(sys.argv[2]) if len(sys.argv) > 2 else 1
preset = sys.argv[3] if len(sys.argv) > 3 else "poisson"
cfg = S.Config.poisson_saturne() if preset == "poisson" else S.Config.solar_sail()
cfg.width = cfg.height = 2048
sms = torch.cuda.get_device_properties(0).multi_processor_count
lanes = sms * 128 * wps
pod = cfg.to_pod()
pod.iterations = 1_000_000_000 // lanes
N.check(L.sar_set_option(b"defer", defer))
rt = C.c_void_p()
N.check(L.sar_runtime_new(2048, 2048, 0, C.byref(rt)))
for _ in range(3):
    N.check(L.sar_runtime_reset_async(rt, None))
    N.check(L.sar_render_seeded_async(C.byref(pod), rt, 1234, 0, lanes, lanes, None))
    N.check(L.sar_runtime_max_async(rt, 0, 0, None))
    N.check(L.sar_colorize_rows_async(C.byref(pod), rt, 0, 0, None, None))
    N.check(L.sar_stream_synchronize(rt, None))
print("done", lanes, pod.iterations)
