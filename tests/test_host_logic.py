"""Host-side logic of the multi-GPU driver (pure functions) and of the oracle's own
render_parallel: decomposition arithmetic of lib.rs:1056-1062 and merge semantics."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def D():
    from strange_attractor_renderer_b200 import dist

    return dist


def test_iterations_per_job_is_two_integer_divisions(D):
    # lib.rs:1058: iterations / num_threads / jobs_per_thread
    assert D.iterations_per_job(1_000_000_000, 132_608, 1) == 7_541
    assert D.iterations_per_job(10_000_123, 96, 3) == 10_000_123 // 96 // 3 == 34_722
    assert D.iterations_per_job(100, 7, 3) == 4            # (100//7)//3, not 100//21 rounded differently
    for world in (1, 2, 4, 8):                              # weak scaling keeps the per-job length
        assert D.iterations_per_job(world * 10**9, world * 132_608, 1) == 7_541


def test_job_slices_partition_the_job_list(D):
    for world in (1, 2, 3, 8):
        lanes, jpt = 96, 5
        seen = []
        for r in range(world):
            first, n = D.job_slice(r, world, lanes, jpt)
            seen.extend(range(first, first + n))
        assert seen == list(range(world * lanes * jpt))    # lib.rs:1062: jobs_per_thread * num_threads jobs


def test_stripes_partition_the_rows(D):
    for world in (1, 2, 3, 8):
        for h in (1, 7, 2000, 2048, 4096):
            rows = []
            for r in range(world):
                r0, n = D.stripe_rows(r, world, h)
                rows.extend(range(r0, r0 + n))
            assert rows == list(range(h))


def test_oracle_merge_in_job_order_equals_sequential(oracle):
    """Why the GPU merge is deterministic: merging per-worker Runtimes in job order with 'ties keep
    self' (lib.rs:728) is exactly one Runtime rendered job after job."""
    cfg = oracle.poisson_saturne()
    cfg.width, cfg.height, cfg.iterations = 160, 120, 8000
    pts = oracle.seed_points(9, 0, 24)
    seq = oracle.Runtime(160, 120)
    oracle.render_jobs(cfg, seq, pts)
    parts = []
    for k in range(3):
        rt = oracle.Runtime(160, 120)
        oracle.render_jobs(cfg, rt, pts[8 * k:8 * k + 8])
        parts.append(rt)
    parts[0].merge(parts[1])
    parts[0].merge(parts[2])
    assert np.array_equal(parts[0].count, seq.count) and parts[0].max == seq.max
    assert np.array_equal(parts[0].zbuf, seq.zbuf)
    assert np.array_equal(parts[0].steps, seq.steps)
    with pytest.raises(ValueError):
        parts[0].merge(oracle.Runtime(10, 10))            # assert_eq! panic, lib.rs:709-710


def test_oracle_render_parallel_counts_match_sequential(oracle):
    """The threaded CPU baseline (dynamic job counter, lib.rs:962-982) computes the same counts as
    the sequential semantics; only exact z ties may differ with scheduling."""
    cfg = oracle.solar_sail()
    cfg.width, cfg.height, cfg.iterations = 180, 200, 4 * 6 * 5000 + 17
    pts = oracle.seed_points(3, 0, 24)
    img, merged = oracle.render_parallel(cfg, 4, 6, pts, want_runtime=True)
    seq_cfg = cfg.copy()
    seq_cfg.iterations = cfg.iterations // 4 // 6
    assert seq_cfg.iterations == 5000
    seq = oracle.Runtime(180, 200)
    oracle.render_jobs(seq_cfg, seq, pts)
    assert np.array_equal(merged.count, seq.count) and merged.max == seq.max
    assert np.array_equal(merged.zbuf, seq.zbuf)
    assert (merged.steps != seq.steps).mean() < 1e-3
    assert img.shape == (200, 180, 4)


def test_angle_iter_follows_the_reference_cli():
    """AngleIter (src/bin/main.rs:107-176): while curr + step/2 < end yield curr in RADIANS
    (main.rs:166), curr += step; if nothing was yielded, the start value as is (main.rs:169-171)."""
    import math

    import strange_attractor_renderer_b200 as S

    sweep = S.angle_iter(0.0, 360.0, 1.0)                      # BASELINE configs[4]: 360 frames
    assert len(sweep) == 360 and sweep[0] == 0.0
    assert all(sweep[k] == float(k) * math.pi / 180.0 for k in range(360))
    assert S.angle_iter(10.0, 20.0, 4.0) == [a * math.pi / 180.0 for a in (10.0, 14.0)]   # 18 + 2 >= 20 stops
    assert S.angle_iter(220.0, 220.0, 1.0) == [220.0]          # single image: passed through unconverted
    assert S.angle_iter(5.0, 5.4, 1.0) == [5.0]


def test_product_library_has_no_diagnostic_kernels():
    """no switch in the product library may make a handle produce incomplete
    results.  diagnostic_mode != 0 is refused, and only the product instantiations of the iterate
    kernel (MODE 0) are in the binary."""
    import subprocess

    from strange_attractor_renderer_b200 import _native as N

    L = N.lib()
    assert L.sar_set_option(b"diagnostic_mode", 0) == 0
    for m in (1, 2, 4, 7):
        assert L.sar_set_option(b"diagnostic_mode", m) == N.SAR_ERR_UNSUPPORTED
    assert L.sar_set_option(b"defer", 1) == N.SAR_ERR_INVALID      # removed knob
    for nt in (1, 2, 4):
        assert L.sar_set_option(b"traj_per_thread", nt) == 0
    assert L.sar_set_option(b"traj_per_thread", 3) == N.SAR_ERR_INVALID
    assert L.sar_set_option(b"traj_per_thread", 1) == 0
    assert L.sar_set_option(b"pipeline", 1) == 0 and L.sar_set_option(b"pipeline", 2) == N.SAR_ERR_INVALID
    assert L.sar_set_option(b"pipeline", 0) == 0
    assert L.sar_set_option(b"tile_scatter", 0) == 0 and L.sar_set_option(b"tile_scatter", 2) == N.SAR_ERR_INVALID
    assert L.sar_set_option(b"tile_scatter", 1) == 0
    syms = subprocess.run(["cuobjdump", "-symbols", N.LIB_PATH], capture_output=True, text=True).stdout
    kernels = sorted({w for line in syms.splitlines() if "STO_ENTRY" in line for w in line.split() if "iterate_kernel" in w})
    assert kernels and all("ELi0ELi" in k for k in kernels)   # iterate_kernel<NT, MODE = 0, PIPE>, kernels


def test_angle_iter_files_follow_the_binary_s_naming():
    """AngleIter (src/bin/main.rs:106-176): degrees -> radians, zero-padded frame index sized by the estimated count, the
    single-image branch for an empty range (angle left unconverted), and its quirks (no digits for <= 1.5 steps;
    set_extension replacing a dotted stem's tail)."""
    import math

    import strange_attractor_renderer_b200 as S

    seq = S.angle_iter_files(0.0, 360.0, 1.0, "out/attractor.png")
    assert len(seq) == 360 and seq[0] == (0.0, "out/attractor000.png") and seq[359][1] == "out/attractor359.png"
    assert seq[7][0] == 7.0 * math.pi / 180.0
    assert [a for a, _ in seq] == S.angle_iter(0.0, 360.0, 1.0)
    assert [p for _, p in S.angle_iter_files(0.0, 100.0, 10.0, "a.bmp")] == [f"a{i}.bmp" for i in range(10)]       # count 9.5 -> 1 digit
    assert [p for _, p in S.angle_iter_files(0.0, 10.0, 5.0, "x.png")] == ["x.png", "x.png"]                        # count 1.5 -> no digits
    assert S.angle_iter_files(90.0, 90.0, 1.0, "one.png") == [(90.0, "one.png")]                                    # single image, radians as given
    assert S.angle_iter_files(0.0, 30.0, 1.0, "noext")[3][1] == "noext03"
    assert S.angle_iter_files(0.0, 30.0, 1.0, "a.b.png")[3][1] == "a.png"                                           # "a.b03" -> set_extension("png")
