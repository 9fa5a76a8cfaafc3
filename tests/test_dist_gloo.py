"""The N > 1 host path on CPU: world_size-2 gloo processes exercise the torch.distributed plumbing
(handle all-gather, max all-reduce, barriers) and the job/stripe partition; the per-rank render is
stood in for by the CPU oracle, so the test also shows that trajectory sharding + rank-ordered
merge reproduces the sequential result exactly."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank: int, world: int, port: int, out_dir: str):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    import torch.distributed as dist

    from oracle import oracle as O
    from strange_attractor_renderer_b200 import dist as D

    group = D.init_process_group(world, rank, 0, backend="gloo")
    # plumbing
    handles = D.allgather_bytes(bytes([rank]) * 64, group, 0)
    assert handles == [bytes([r]) * 64 for r in range(world)]
    assert D.allreduce_max_u32(1000 + rank, group, 0) == 1000 + world - 1
    assert D.max_over_ranks(1.5 * (rank + 1), group, 0) == 1.5 * world
    D.barrier(group)
    # a frame: rank r renders its job slice, then stripes are reduced in rank (= job) order
    lanes, jpt, total = 6, 2, 2 * 6 * 2 * 3000 + 5
    cfg = O.solar_sail()
    cfg.width, cfg.height = 90, 101
    cfg.iterations = D.iterations_per_job(total, lanes * world, jpt)
    first, n = D.job_slice(rank, world, lanes, jpt)
    rt = O.Runtime(90, 101)
    O.render_jobs(cfg, rt, O.seed_points(77, first, n))
    gathered = [None] * world
    dist.all_gather_object(gathered, (rt.count.copy(), rt.steps.copy(), rt.zbuf.copy()))
    row0, rows = D.stripe_rows(rank, world, 101)
    acc = O.Runtime(90, 101)
    for (c, s, z) in gathered:                         # rank order == job order: ties keep the earlier job
        part = O.Runtime(90, 101)
        part.load(c, s, z)
        acc.merge(part)
    stripe_max = int(acc.count[row0:row0 + rows].max())
    gmax = D.allreduce_max_u32(stripe_max, group, 0)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), row0=row0, rows=rows, count=acc.count[row0:row0 + rows],
             steps=acc.steps[row0:row0 + rows], zbuf=acc.zbuf[row0:row0 + rows], gmax=gmax)
    D.barrier(group)
    D.shutdown(group)


def test_two_rank_frame_on_gloo(tmp_path, oracle):
    import torch.multiprocessing as mp

    from strange_attractor_renderer_b200 import dist as D

    world, port = 2, 29551
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    lanes, jpt, total = 6, 2, 2 * 6 * 2 * 3000 + 5
    cfg = oracle.solar_sail()
    cfg.width, cfg.height = 90, 101
    cfg.iterations = D.iterations_per_job(total, lanes * world, jpt)
    assert cfg.iterations == 3000
    seq = oracle.Runtime(90, 101)
    oracle.render_jobs(cfg, seq, oracle.seed_points(77, 0, lanes * jpt * world))
    count = np.zeros_like(seq.count); steps = np.zeros_like(seq.steps); zbuf = np.zeros_like(seq.zbuf)
    covered = 0
    for r in range(world):
        d = np.load(tmp_path / f"rank{r}.npz")
        a, m = int(d["row0"]), int(d["rows"])
        count[a:a + m], steps[a:a + m], zbuf[a:a + m] = d["count"], d["steps"], d["zbuf"]
        covered += m
        assert int(d["gmax"]) == seq.max                  # Runtime.max is global (lib.rs:860)
    assert covered == 101
    assert np.array_equal(count, seq.count)
    assert np.array_equal(zbuf, seq.zbuf)
    assert np.array_equal(steps, seq.steps)
