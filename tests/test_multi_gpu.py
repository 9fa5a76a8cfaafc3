"""N-rank frame == oracle, bit for bit (needs >= 2 GPUs; skipped otherwise — bench.py --gpus N runs the
same check inside the driver's scaling run and reports it as "parity")."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("preset,kind,world", [("solar", "gas", 2), ("poisson", "depth", 2), ("solar", "depth", 4), ("poisson", "gas", 8)])
def test_n_rank_frame_equals_oracle(preset, kind, world):
    """N-rank frame (incl. RenderKind::Depth, whose min/max is global) vs the oracle, bit for bit."""
    if _gpu_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29533 + world), os.path.join(ROOT, "tools", "dist_check.py"), preset, kind]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "DIST_CHECK_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_baseline_cfg3_full_size_8_ranks_equals_oracle():
    """BASELINE configs[3] itself: poisson-saturne, 8e9 iterations, 4096x4096, 8 ranks row-striped — count, zbuf, steps and
    image against the oracle (about a minute of host time for the oracle's 9e9 steps)."""
    if _gpu_count() < 8:
        pytest.skip("needs 8 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "8", "--master-addr", "127.0.0.1",
           "--master-port", "29547", os.path.join(ROOT, "tools", "dist_check.py"), "poisson", "gas", "full"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0 and "DIST_CHECK_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_in_process_two_device_renderer_equals_one_device():
    """ParallelRenderer over two devices in ONE process (frames' jobs split by device, merged by
    peer copy) gives the image and buffers of one device running the same num_threads."""
    if _gpu_count() < 2:
        pytest.skip("needs 2 GPUs")
    import numpy as np

    import strange_attractor_renderer_b200 as S

    cfg = S.Config.solar_sail()
    cfg.width, cfg.height, cfg.iterations, cfg.angle = 300, 333, 20_000_000, 0.7
    two = S.ParallelRenderer.new(devices=[0, 1], threads=512)
    one = S.ParallelRenderer.new(devices=[0], threads=1024)
    assert two.num_threads() == one.num_threads() == 1024
    a = S.render_parallel(two, cfg, 3, seed=31)
    b = S.render_parallel(one, cfg, 3, seed=31)
    assert np.array_equal(a, b)
    sa, sb = two.runtime().download(), one.runtime().download()
    for x, y in zip(sa[:3], sb[:3]):
        assert np.array_equal(x.view(np.uint8), y.view(np.uint8))
    # a sequence round-robins frames over the devices; frames do not depend on which device ran them
    angles = S.angle_iter(0.0, 40.0, 10.0)
    fa = S.render_sequence(two, cfg, angles, 2, seed=9, shared_points=True)
    fb = S.render_sequence(S.ParallelRenderer.new(devices=[1], threads=512), cfg, angles, 2, seed=9, shared_points=True)
    assert np.array_equal(fa, fb)
    two.shutdown()
    one.shutdown()
