"""Multi-GPU parity check, run under torch.distributed.run (one process per GPU): the N-rank frame
(trajectory-sharded render + NVLink stripe merge + stripe colourise into rank 0) must equal the CPU
oracle on the same job list, bit for bit — count, zbuf, steps and the RGBA16 image.
Usage: dist_check.py [solar|poisson] [gas|depth] [full].  `full` = BASELINE configs[3] at full size (poisson-saturne, 1e9 iterations
per GPU, 4096x4096, default lanes; the oracle side takes about a minute on 16 cores).  Prints DIST_CHECK_OK on rank 0."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from bench import parity_frame
from strange_attractor_renderer_b200 import dist as D

world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
group = D.init_process_group(world, rank, local)
preset = sys.argv[1] if len(sys.argv) > 1 else "solar"
depth = len(sys.argv) > 2 and sys.argv[2] == "depth"
if len(sys.argv) > 3 and sys.argv[3] == "full":
    checks = parity_frame(world, rank, local, group, preset, depth=depth, per_gpu=1_000_000_000, lanes=0, jpt=1, width=4096, height=4096, seed=1234)
else:
    checks = parity_frame(world, rank, local, group, preset, depth=depth)
ok = True
if rank == 0:
    print("checks", checks, flush=True)
    ok = all(checks[k] for k in ("count", "zbuf", "steps", "image")) and checks["recorded"] > 0
flag = torch.tensor([1 if ok else 0], device=f"cuda:{local}")
torch.distributed.broadcast(flag, 0)
D.shutdown(group)
if rank == 0:
    print("DIST_CHECK_OK" if ok else "DIST_CHECK_FAILED", flush=True)
sys.exit(0 if int(flag.item()) == 1 else 1)
