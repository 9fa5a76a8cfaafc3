"""Multi-GPU parity check, run under torch.distributed.run (one process per GPU):
the N-rank frame (trajectory-sharded render + NVLink stripe merge + stripe colourise into rank 0)
must be bit-identical to the same job list rendered on one GPU.  Prints DIST_CHECK_OK on rank 0."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import strange_attractor_renderer_b200 as S
from strange_attractor_renderer_b200 import _native as N
from strange_attractor_renderer_b200 import dist as D

world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
group = D.init_process_group(world, rank, local)
L = N.lib()

preset = sys.argv[1] if len(sys.argv) > 1 else "solar"
cfg = S.Config.solar_sail() if preset == "solar" else S.Config.poisson_saturne()
cfg.width, cfg.height, cfg.angle = 450, 501, 1.25       # height not divisible by the world size
lanes, jpt, per_gpu = 2048, 2, 40_000_000
frame = D.Frame(cfg, device=local, world=world, rank=rank, group=group, lanes=lanes, jobs_per_thread=jpt,
                iterations_per_gpu=per_gpu, seed=4321)
stream = torch.cuda.Stream(device=local)
sp = C.c_void_p(stream.cuda_stream)
for rep in range(2):                                      # twice: reset/re-merge must be clean
    frame.step_device(sp)
torch.cuda.synchronize()
frame.check_sync()
D.barrier(group)

# every rank downloads its merged stripe; rank 0 assembles the full state
count = np.empty((frame.h, frame.w), np.uint32); steps = np.empty((frame.h, frame.w), np.float64); zbuf = np.empty((frame.h, frame.w), np.float32)
N.check(L.sar_runtime_download(frame.rt, count.ctypes.data_as(N._u32p), steps.ctypes.data_as(N._f64p), zbuf.ctypes.data_as(N._f32p), None))
r0, n = frame.row0, frame.rows
parts = [None] * world
dist.all_gather_object(parts, (r0, n, count[r0:r0 + n].copy(), steps[r0:r0 + n].copy(), zbuf[r0:r0 + n].copy()))
ok = True
if rank == 0:
    img = np.empty((frame.h, frame.w, 4), np.uint16)
    N.check(L.sar_runtime_image_download(frame.rt, 0, 0, img.ctypes.data_as(N._u16p), None))
    for (a, m, c, s, z) in parts:
        count[a:a + m], steps[a:a + m], zbuf[a:a + m] = c, s, z
    # the same job list on ONE GPU
    one = cfg.to_pod()
    one.iterations = frame.iterations_per_job
    rt = C.c_void_p()
    N.check(L.sar_runtime_new(frame.w, frame.h, local, C.byref(rt)))
    N.check(L.sar_render_seeded_async(C.byref(one), rt, 4321, 0, lanes * jpt * world, lanes * world, None))
    c1 = np.empty_like(count); s1 = np.empty_like(steps); z1 = np.empty_like(zbuf); mx = C.c_uint32()
    N.check(L.sar_runtime_download(rt, c1.ctypes.data_as(N._u32p), s1.ctypes.data_as(N._f64p), z1.ctypes.data_as(N._f32p), C.byref(mx)))
    img1 = np.empty_like(img)
    N.check(L.sar_colorize(C.byref(one), rt, img1.ctypes.data_as(N._u16p), None))
    checks = {
        "count": np.array_equal(count, c1), "zbuf": np.array_equal(zbuf.view(np.uint32), z1.view(np.uint32)),
        "steps": np.array_equal(steps.view(np.uint64), s1.view(np.uint64)), "image": np.array_equal(img, img1),
        "nonempty": int(c1.sum()) > 0,
    }
    ok = all(checks.values())
    print("checks", checks, "recorded", int(c1.sum(dtype=np.uint64)), "max", mx.value, flush=True)
    L.sar_runtime_free(rt)
flag = torch.tensor([1 if ok else 0], device=f"cuda:{local}")
dist.broadcast(flag, 0)
frame.close()
D.shutdown(group)
if rank == 0:
    print("DIST_CHECK_OK" if ok else "DIST_CHECK_FAILED", flush=True)
sys.exit(0 if int(flag.item()) == 1 else 1)
