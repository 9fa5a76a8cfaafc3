"""The multithreaded mode of the oracle (orc_render_jobs_mt) is the serial job loop, bit for bit.

It exists so that the GPU can be checked against the oracle at the BASELINE sizes (1e9
iterations); this file pins it to the plain serial loop (orc_render_jobs = n reference render()
calls in list order, lib.rs:742-747) on cases that exercise what could go wrong: z ties across
thread slices (a start point repeated in two slices ties on every hit), NaN trajectories, a Runtime that
already holds earlier renders, more threads than jobs."""
import numpy as np
import pytest


def _same(a, b):
    assert np.array_equal(a.count, b.count)
    assert np.array_equal(a.zbuf.view(np.uint32), b.zbuf.view(np.uint32))
    assert np.array_equal(a.steps.view(np.uint64), b.steps.view(np.uint64))
    assert a.max == b.max


@pytest.mark.parametrize("preset,w,h,iters,jobs,threads", [
    ("poisson", 24, 20, 30_000, 37, 5),      # ~2000 hits per pixel
    ("solar", 90, 100, 4_000, 64, 8),        # 38 % of the start points diverge to NaN
    ("poisson", 64, 64, 2_000, 3, 16),       # more threads than jobs
    ("solar", 40, 30, 10_000, 50, 7),
])
def test_mt_equals_serial(oracle, preset, w, h, iters, jobs, threads):
    cfg = oracle.poisson_saturne() if preset == "poisson" else oracle.solar_sail()
    cfg.width, cfg.height, cfg.iterations = w, h, iters
    if preset == "solar":
        cfg.angle = 3.839724354387525
    pts = oracle.seed_points(77, 0, jobs)
    a, b = oracle.Runtime(w, h), oracle.Runtime(w, h)
    sa, sb = oracle.OrcStats(), oracle.OrcStats()
    oracle.render_jobs(cfg, a, pts, sa)
    oracle.render_jobs_mt(cfg, b, pts, threads, sb)
    _same(a, b)
    assert sa.recorded == sb.recorded and sa.nan_iters == sb.nan_iters


def test_mt_continues_a_non_reset_runtime(oracle):
    """render() accumulates into a non-reset Runtime (lib.rs:742-743): thread 0 of the MT mode works
    in place, so earlier renders keep every tie."""
    cfg = oracle.poisson_saturne()
    cfg.width, cfg.height, cfg.iterations = 30, 30, 20_000
    first, second = oracle.seed_points(5, 0, 9), oracle.seed_points(6, 0, 21)
    a, b = oracle.Runtime(30, 30), oracle.Runtime(30, 30)
    oracle.render_jobs(cfg, a, np.concatenate([first, second]))
    oracle.render_jobs(cfg, b, first)
    oracle.render_jobs_mt(cfg, b, second, 6)
    _same(a, b)


def test_mt_ties_between_slices_keep_the_earlier_job(oracle):
    """Exact z ties between different trajectories are vanishingly rare in f32, so they are forced:
    the same start points appear in two thread slices (every hit of the copy ties), and a hand-made
    pair of Runtimes checks that the merge the MT mode uses keeps `self` on a tie (lib.rs:728)."""
    cfg = oracle.poisson_saturne()
    cfg.width, cfg.height, cfg.iterations = 48, 40, 5_000
    base = oracle.seed_points(9, 0, 6)
    pts = np.concatenate([base, oracle.seed_points(10, 0, 5), base])
    a, b = oracle.Runtime(48, 40), oracle.Runtime(48, 40)
    sa = oracle.OrcStats()
    oracle.render_jobs(cfg, a, pts, sa)
    oracle.render_jobs_mt(cfg, b, pts, 4)
    assert sa.z_ties > 0
    _same(a, b)
    x, y = oracle.Runtime(2, 1), oracle.Runtime(2, 1)
    x.load(np.array([[1, 1]], np.uint32), np.array([[0.25, 0.25]]), np.array([[0.5, 0.5]], np.float32))
    y.load(np.array([[2, 2]], np.uint32), np.array([[0.75, 0.75]]), np.array([[0.5, 0.6]], np.float32))
    x.merge(y)
    assert x.count.tolist() == [[3, 3]] and x.steps.tolist() == [[0.25, 0.75]] and x.max == 3
